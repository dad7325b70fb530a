"""Run one of the reference's scripts UNCHANGED with the B200 generator behind ``import model``.

    python -m sr_gan_fd_b200.compat.run ESRGAN inference.py --device_type cuda --inputs_path figure/baboon_lr.png ...
    python -m sr_gan_fd_b200.compat.run --set rrdbnet_config.epochs=1 --set rrdbnet_config.batch_size=4 ESRGAN train_rrdbnet.py

What it does (nothing in the reference tree is edited or copied):
  * registers the folder's compat shim (``compat/esrgan_model.py`` ...) as the module named ``model``, so the script's own
    ``import model`` / ``model.__dict__[arch](**kw)`` (ESRGAN/train_rrdbnet.py:174-178, inference.py:41-45) resolves to it;
  * puts the reference folder on ``sys.path`` for the script's sibling imports (``imgproc``, ``dataset``, ``utils``, ``*_config``);
  * optional ``--set module.attr=value`` edits a config module IN MEMORY before the script imports it (the reference keeps its
    run configuration in ``*_config.py`` files users edit by hand: dataset paths, epochs, batch size);
  * optional ``--vgg`` swaps the folder's ``ContentLoss`` for the B200 VGG19 feature path (``sr_gan_fd_b200.vgg``);
  * optional ``--disc`` swaps ``DiscriminatorUNet`` / ``discriminator_unet`` (BSRGAN, Real_ESRGAN) for the B200 U-Net discriminator
    (``sr_gan_fd_b200.discriminator``);
  * optional ``--iqa`` also registers a shim for ``image_quality_assessment`` whose ``PSNR`` / ``SSIM`` are the fused B200 kernels;
  * optional ``--stub name`` installs an empty stand-in for a logging dependency that is not installed (mlflow, lpips, ...);
  * the script then runs under ``runpy`` with ``__name__ == "__main__"`` from the current working directory.
"""
from __future__ import annotations

import argparse
import ast
import importlib
import os
import runpy
import sys
import types

SHIMS = {"ESRGAN": "esrgan_model", "BSRGAN": "bsrgan_model", "Real_ESRGAN": "real_esrgan_model", "A-ESRGAN": "a_esrgan_model"}


class _Anything(types.ModuleType):
    """Stand-in module: every attribute is a callable / context manager that does nothing."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Noop()


class _Noop:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default=None)
    ap.add_argument("--set", action="append", default=[], metavar="module.attr=value")
    ap.add_argument("--stub", action="append", default=[], metavar="module")
    ap.add_argument("--stock", action="store_true", help="do NOT install the shim: run the reference's own model.py (A/B runs)")
    ap.add_argument("--iqa", action="store_true", help="also replace image_quality_assessment.PSNR / SSIM by the fused B200 versions")
    ap.add_argument("--vgg", action="store_true", help="also replace model.ContentLoss / content_loss by the B200 VGG19 feature path")
    ap.add_argument("--disc", action="store_true", help="also replace model.DiscriminatorUNet / discriminator_unet by the B200 U-Net discriminator")
    ap.add_argument("folder", choices=sorted(SHIMS))
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)

    if args.reference is None:
        from ._passthrough import reference_root
        args.reference = reference_root()
    os.environ["SRGANFD_REFERENCE"] = args.reference
    folder = os.path.join(args.reference, args.folder)
    script = os.path.join(folder, args.script)
    if not os.path.isfile(script):
        raise SystemExit(f"{script} not found")
    sys.dont_write_bytecode = True
    for name in args.stub:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            sys.modules.setdefault(".".join(parts[:i]), _Anything(".".join(parts[:i])))
    if folder not in sys.path:
        sys.path.insert(0, folder)
    if not args.stock:
        shim = importlib.import_module(f"sr_gan_fd_b200.compat.{SHIMS[args.folder]}")
        if args.vgg:
            from sr_gan_fd_b200 import vgg as _vgg
            cls = _vgg.ContentLoss if args.folder == "ESRGAN" else _vgg.ContentLossMulti
            ns = dict(shim.__dict__)
            ns["ContentLoss"] = cls
            ns["content_loss"] = lambda **kwargs: cls(**kwargs)
            shim = types.ModuleType("model")
            shim.__dict__.update(ns)
        if args.disc and args.folder in ("BSRGAN", "Real_ESRGAN"):
            from sr_gan_fd_b200 import discriminator as _disc
            ns = dict(shim.__dict__)
            ns["DiscriminatorUNet"] = _disc.DiscriminatorUNet
            ns["discriminator_unet"] = _disc.discriminator_unet
            shim = types.ModuleType("model")
            shim.__dict__.update(ns)
        sys.modules["model"] = shim
        if args.iqa:
            from sr_gan_fd_b200.compat import iqa_module
            sys.modules["image_quality_assessment"] = iqa_module.make(args.folder)
    for item in args.set:
        target, _, value = item.partition("=")
        mod_name, _, attr = target.rpartition(".")
        mod = importlib.import_module(mod_name)
        try:
            val = ast.literal_eval(value)
        except (ValueError, SyntaxError):
            val = value
        if attr == "device":
            import torch
            val = torch.device(val)
        setattr(mod, attr, val)
    sys.argv = [script] + [a for a in args.script_args if a != "--"]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
