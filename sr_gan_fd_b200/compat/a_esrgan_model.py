"""Drop-in for ``A-ESRGAN/model.py``: the RRDB generator entry point ``bsrgan_x2`` / ``BSRGAN`` (A-ESRGAN/model.py:479-558, a
copy of BSRGAN/model.py's) runs on the B200 path; everything else the folder defines -- the attention U-Net critics
(``uNetDiscriminatorAesrgan``), ``ContentLoss``, and the other generator variants (``Generator_RRDB`` / ``gen_rrdb2x``,
``Generator_RPA``, ``BSRGANsa``, ``BSRGANtrans`` / ``bsrgantrans_x2`` with its transformer bottleneck) -- is passed through
as the reference's own torch modules (out of scope).  ``basicsr`` is only needed for a registry decorator; a
no-op stand-in is installed when the package is absent."""
from ..rrdbnet import BSRGAN, bsrgan_x2
from ._passthrough import export as _export

_export(globals(), "A-ESRGAN", dict(BSRGAN=BSRGAN, bsrgan_x2=bsrgan_x2))
