"""Drop-in for a folder's ``image_quality_assessment.py``: the fused B200 ``PSNR`` / ``SSIM`` (``sr_gan_fd_b200.iqa``) + everything
else the reference module defines (NIQE, the numpy variants, ...) passed through.  ``make(folder)`` builds the module object
that ``python -m sr_gan_fd_b200.compat.run --iqa`` registers as ``image_quality_assessment``."""
import types

from ..iqa import PSNR, SSIM
from ._passthrough import export as _export


def make(folder: str) -> types.ModuleType:
    mod = types.ModuleType("image_quality_assessment")
    _export(mod.__dict__, folder, dict(PSNR=PSNR, SSIM=SSIM), module="image_quality_assessment")
    return mod
