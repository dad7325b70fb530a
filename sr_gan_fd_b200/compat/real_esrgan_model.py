"""Drop-in for ``Real_ESRGAN/model.py``: positional-ctor RRDBNet with the pixel-unshuffle front."""
from typing import Any

from ..rrdbnet import RealRRDBNet as RRDBNet
from ._passthrough import export as _export


def rrdbnet_x4(**kwargs: Any) -> RRDBNet:
    return RRDBNet(upscale_factor=4, **kwargs)


_export(globals(), "Real_ESRGAN", dict(RRDBNet=RRDBNet, rrdbnet_x4=rrdbnet_x4))
