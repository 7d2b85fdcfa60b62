"""The C-ABI library loads on a CPU-only box and exports every symbol include/wm_b200.h declares, with the
argument counts the ctypes binding uses.  No compute calls (no GPU here)."""
import os
import re

import pytest

from wildlifemapper_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_decls():
    src = open(os.path.join(ROOT, "include", "wm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:int|const char\*)\s+(wm_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args == "void" else len([a for a in args.split(",") if a.strip()])
    return decls


def test_header_symbols_exported_and_bound():
    decls = header_decls()
    assert len(decls) >= 17
    so = lib.load()
    for name, nargs in decls.items():
        assert hasattr(so, name), f"{name} declared in the header but not exported"
        assert name in lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(lib.SIGNATURES[name]) == nargs, (name, nargs, len(lib.SIGNATURES[name]))
    assert set(lib.SIGNATURES) == set(decls)


def test_no_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    so = lib.load()
    assert so.wm_version() >= 100
    rc = so.wm_device_check()
    assert rc == -3  # WM_ERR_ARCH
    assert b"CUDA" in so.wm_last_error() or b"sm_" in so.wm_last_error()
    with pytest.raises(lib.WmError):
        lib.call("wm_transpose", None, None, 1, 4, 4, 2, None)


def test_product_never_imports_oracle():
    """The shipped package must not route through the CPU oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "wildlifemapper_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
