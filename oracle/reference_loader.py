"""ORACLE helper (test infrastructure): import the UNMODIFIED reference ``model.py`` / IQA modules from
``/root/reference`` when that tree is present (build container only -- it never exists on the GPU box).

Nothing is copied: the modules are executed where they lie, with ``sys.dont_write_bytecode`` so the read-only tree
is not touched.  ``A-ESRGAN/model.py`` imports ``basicsr`` at import time only (line 24, 30); a 3-line stub
registry is installed for it.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SRGANFD_REFERENCE", "/root/reference")
FOLDERS = {"esrgan": "ESRGAN", "bsrgan": "BSRGAN", "real": "Real_ESRGAN", "aesrgan": "A-ESRGAN"}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ESRGAN", "model.py"))


def _install_basicsr_stub() -> None:
    if "basicsr" in sys.modules:
        return

    class _Registry:
        def register(self, *a, **k):
            return (lambda f: f) if not a or not callable(a[0]) else a[0]

    basicsr = types.ModuleType("basicsr")
    utils = types.ModuleType("basicsr.utils")
    registry = types.ModuleType("basicsr.utils.registry")
    registry.ARCH_REGISTRY = _Registry()
    basicsr.utils = utils
    utils.registry = registry
    sys.modules.update({"basicsr": basicsr, "basicsr.utils": utils, "basicsr.utils.registry": registry})


def load_module(flavour: str, module: str = "model"):
    """Import ``<folder>/<module>.py`` from the reference under a private name (no sys.modules collision between
    the four same-named ``model`` modules)."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    folder = os.path.join(REFERENCE_ROOT, FOLDERS[flavour])
    sys.dont_write_bytecode = True
    if flavour == "aesrgan":
        _install_basicsr_stub()
    private = f"_srganfd_ref_{flavour}_{module}"
    if private in sys.modules:
        return sys.modules[private]
    sys.path.insert(0, folder)  # the reference modules import their siblings (imgproc, *_config) by bare name
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("model", "imgproc", "image_quality_assessment", "aesrgan_config", "utils")}
    try:
        spec = importlib.util.spec_from_file_location(private, os.path.join(folder, module + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[private] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(folder)
        for k in ("model", "imgproc", "image_quality_assessment", "aesrgan_config", "utils"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return mod


def build_generator(flavour: str, upscale_factor: int = 4, num_blocks: int = 23, in_channels: int = 3,
                    out_channels: int = 3, channels: int = 64, growth: int = 32):
    """Instantiate the reference generator class of one folder with explicit sizes."""
    m = load_module(flavour)
    if flavour == "esrgan":
        return m.RRDBNet(in_channels, out_channels, channels, growth, num_blocks, upscale_factor)
    if flavour in ("bsrgan", "aesrgan"):
        return m.BSRGAN(in_channels, out_channels, channels, growth, num_blocks, upscale_factor)
    if flavour == "real":
        return m.RRDBNet(in_channels, out_channels, channels, growth, num_blocks, upscale_factor)
    raise ValueError(flavour)
