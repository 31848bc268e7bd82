"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/osb.h declares,
the ctypes signatures agree with the header's parameter counts, and nothing computes without a GPU."""
import os
import re

import pytest
import torch

from openset_imagenet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "osb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(osb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


def test_header_and_binding_agree():
    protos = _header_prototypes()
    assert len(protos) >= 20
    assert set(protos) == set(_lib.SIGNATURES), set(protos) ^ set(_lib.SIGNATURES)
    for name, n in protos.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name


def test_library_exports_every_symbol():
    lib = _lib.load()
    for name in _header_prototypes():
        assert hasattr(lib, name), name
    assert lib.osb_abi_version() == _lib.ABI_VERSION


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import numpy as np
    import openset_imagenet_b200 as osb
    lib = _lib.load()
    assert lib.osb_device_check() == _lib.ERR_ARCH
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        osb.EntropicOpensetLoss(4)(torch.zeros(2, 4), torch.tensor([0, -1]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        osb.confidence(torch.rand(3, 4), torch.tensor([0, 1, -1]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        osb.calculate_oscr(np.array([0, -1]), np.random.rand(2, 3).astype(np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "openset_imagenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)
