"""Drop-in for ``BSRGAN/model.py`` (and the RRDB path of ``A-ESRGAN/model.py``): B200 generator + reference critics."""
from ..rrdbnet import BSRGAN, bsrgan_x2, bsrgan_x4
from ._passthrough import export as _export

_export(globals(), "BSRGAN", dict(BSRGAN=BSRGAN, bsrgan_x2=bsrgan_x2, bsrgan_x4=bsrgan_x4))
