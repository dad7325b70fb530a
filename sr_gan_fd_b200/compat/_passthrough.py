"""Re-export everything the reference's own model.py defines (discriminators, ContentLoss, ...) so that only the
generator classes / factories are overridden.  The reference file is executed where it lies, never copied."""
import importlib.util
import os
import sys


def _stub_basicsr() -> None:
    """A-ESRGAN/model.py:30 imports ``basicsr.utils.registry.ARCH_REGISTRY`` only to decorate classes; when basicsr is not
    installed a no-op registry with the same ``register()`` surface lets the reference file execute unchanged."""
    try:
        import basicsr.utils.registry  # noqa: F401
        return
    except Exception:
        pass
    import types

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


def reference_root() -> str:
    """Where the reference tree lives: $SRGANFD_REFERENCE, else /root/reference (the build container), else the unmodified copy staged
    under <repo>/baseline/_ref (git-ignored; travels to the GPU box with the snapshot, see __graft_entry__.stage_reference)."""
    env = os.environ.get("SRGANFD_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "baseline", "_ref")


def load_reference_model(folder: str, module: str = "model"):
    root = reference_root()
    path = os.path.join(root, folder, module + ".py")
    if not os.path.isfile(path):
        return None
    name = f"_srganfd_reference_{folder.replace('-', '_').lower()}_{module}"
    if name in sys.modules:
        return sys.modules[name]
    sys.dont_write_bytecode = True
    if folder == "A-ESRGAN":
        _stub_basicsr()
    here = os.path.join(root, folder)
    added = here not in sys.path
    if added:
        sys.path.append(here)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(name, None)
        return None
    return mod


def export(namespace: dict, folder: str, overrides: dict, module: str = "model") -> None:
    ref = load_reference_model(folder, module)
    if ref is not None:
        for k, v in ref.__dict__.items():
            if not k.startswith("__"):
                namespace.setdefault(k, v)
    namespace.update(overrides)
    namespace["__all__"] = sorted(k for k in namespace if not k.startswith("_"))
