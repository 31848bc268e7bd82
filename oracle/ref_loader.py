"""Load the reference's hot-path files UNMODIFIED, from /root/reference (build container) or from the verbatim
copies ``oracle/vendor_ref.py`` placed under ``oracle/_ref/`` (git-ignored; what travels to the GPU box).

TEST INFRASTRUCTURE.  ``import openset_imagenet`` fails here (``robustness``, ``vast`` and
``matplotlib`` are absent), but the files on the path load by file path
once ``vast.tools`` (device placement only, no arithmetic) and ``matplotlib`` (module-level
imports only) are stubbed.  Where neither location exists ``available()`` is False and everything that
needs the reference is skipped; the golden vectors made from it (``oracle/make_golden.py`` ->
``tests/golden/``) always travel.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("OSB_REFERENCE_ROOT", "/root/reference")
_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_cache: dict[str, types.ModuleType] = {}


def package_dir() -> str | None:
    """Directory holding the reference's losses.py / metrics.py / util.py / model.py, or None."""
    for root in (REF_ROOT, _VENDORED):
        pkg = os.path.join(root, "openset_imagenet")
        if os.path.isfile(os.path.join(pkg, "losses.py")):
            return pkg
    return None


def available() -> bool:
    return package_dir() is not None


def source() -> str:
    """'checkout' (/root/reference), 'vendored' (oracle/_ref) or 'absent'."""
    pkg = package_dir()
    return "absent" if pkg is None else ("vendored" if pkg.startswith(_VENDORED) else "checkout")


def _install_stubs() -> None:
    import torch

    if "vast" not in sys.modules:
        vast = types.ModuleType("vast")
        tools = types.ModuleType("vast.tools")
        tools._device = torch.device("cpu")

        def device(x):
            return x.to(tools._device)

        def set_device_cpu():
            tools._device = torch.device("cpu")

        def set_device_gpu(index=0):
            tools._device = torch.device(f"cuda:{index}")

        tools.device, tools.set_device_cpu, tools.set_device_gpu = device, set_device_cpu, set_device_gpu
        vast.tools = tools
        sys.modules["vast"] = vast
        sys.modules["vast.tools"] = tools
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            class _Anything(types.ModuleType):
                def __getattr__(self, name):
                    if name.startswith("__"):
                        raise AttributeError(name)
                    return _Anything(f"{self.__name__}.{name}")

                def __call__(self, *a, **k):
                    return self

            for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.lines",
                         "matplotlib.colors", "matplotlib.cm", "matplotlib.backends",
                         "matplotlib.backends.backend_pdf"):
                sys.modules[name] = _Anything(name)


def load(name: str) -> types.ModuleType:
    """name in {'losses', 'metrics', 'util', 'model'} -> the reference module object."""
    if name in _cache:
        return _cache[name]
    pkg = package_dir()
    if pkg is None:
        raise RuntimeError(f"reference not present at {REF_ROOT} nor vendored under {_VENDORED} (python -m oracle.vendor_ref)")
    _install_stubs()
    spec = importlib.util.spec_from_file_location(f"_osb_ref_{name}", os.path.join(pkg, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod
