"""Drop-in for ``ESRGAN/model.py``: B200 generator + the reference's Discriminator / ContentLoss passed through."""
from ..rrdbnet import RRDBNet, rrdbnet_x1, rrdbnet_x2, rrdbnet_x4, rrdbnet_x8
from ._passthrough import export as _export

_export(globals(), "ESRGAN", dict(RRDBNet=RRDBNet, rrdbnet_x1=rrdbnet_x1, rrdbnet_x2=rrdbnet_x2, rrdbnet_x4=rrdbnet_x4,
                                   rrdbnet_x8=rrdbnet_x8))
