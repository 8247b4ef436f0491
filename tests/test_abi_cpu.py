"""The C-ABI library loads and exports every symbol include/b4d.h declares, the
ctypes mirrors match the C structs, and the product path fails loudly (no CPU
fallback) when there is no CUDA device.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b4d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b4d_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from b4d import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    from b4d import _lib

    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libb4d.so does not export %s" % s
    assert sorted(_lib.EXPORTS) == syms


def test_oracle_exports_the_same_abi(oracle_lib):
    o = oracle_lib.load()
    for s in _declared_symbols():
        if s == "b4d_last_match_stats":
            continue  # device diagnostics only
        assert hasattr(o, s), "liboracle.so does not export %s" % s


def test_struct_mirrors(lib):
    from b4d import _lib

    assert ctypes.sizeof(_lib.Profile) == 56 and ctypes.sizeof(_lib.Stats) == 64
    p = _lib.default_profile()
    assert (p.abi, p.block, p.step, p.search_ht, p.search_wie, p.k_ht, p.k_wie, p.stages) == (1, 4, 3, 11, 11, 16, 32, 2)
    assert abs(p.tau_ht - 2.9527) < 1e-6 and abs(p.tau_wie - 0.7693) < 1e-6 and abs(p.lambda_ht - 2.7) < 1e-6
    assert lib.b4d_version() == 1
    shape = (ctypes.c_int64 * 3)(64, 64, 64)
    assert lib.b4d_num_refs(shape) == 21 ** 3  # 64^3 -> 9 261 reference blocks
    shape = (ctypes.c_int64 * 3)(128, 128, 128)
    assert lib.b4d_num_refs(shape) == 79507


def test_profile_mapping():
    import b4d

    c = b4d.BM4DProfile(search_window_ht=(7, 7, 7), max_stack_size_wiener=16, deterministic=True).to_c(stages=1)
    assert (c.search_ht, c.search_wie, c.k_wie, c.deterministic, c.stages) == (15, 11, 16, 1, 1)
    with pytest.raises(NotImplementedError):
        b4d.BM4DProfile(bs_ht=(8, 8, 8)).to_c()
    with pytest.raises(AttributeError):
        b4d.BM4DProfile(no_such_field=1)


def test_no_cpu_fallback(lib):
    import torch

    import b4d

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        b4d.bm4d(np.zeros((8, 8, 8), np.uint16), 10.0)


def test_call_surface_errors_before_touching_the_device():
    import b4d

    with pytest.raises(NotImplementedError):
        b4d.api._sigma_scalar(np.arange(8.0).reshape(2, 2, 2))
    assert b4d.api._sigma_scalar(np.float32(24)) == 24.0
    with pytest.raises(ValueError):
        b4d.bm4d(np.zeros((8, 8), np.uint16), 10.0)
    with pytest.raises(NotImplementedError):
        b4d.bm4d(np.zeros((8, 8, 8), np.uint16), 10.0, blockmatches=(True, False))
    z, cast = b4d.api._prepare(np.zeros((4, 4, 4), np.int64))
    assert z.dtype == np.uint16 and cast is None
    z, cast = b4d.api._prepare(np.zeros((4, 4, 4), np.float64))
    assert z.dtype == np.float32 and cast == np.float64
    assert b4d.noise_scaled_step(24.0, 0.5) == 12.0 and b4d.noise_scaled_step(1.0, 0.5) == 1.0


def test_bm4d_import_name_shim():
    import bm4d as shim

    import b4d

    assert shim.bm4d is b4d.bm4d and hasattr(shim, "BM4DProfile") and hasattr(shim, "BM4DStages")


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: no file of the package (Python or CUDA/C++) may import,
    include, load or name it, and libb4d.so must not depend on liboracle.so."""
    import os
    import re
    import subprocess

    from b4d import _lib

    pkg = os.path.dirname(os.path.dirname(os.path.abspath(_lib.__file__)))
    pat = re.compile(r"\b(np_oracle|liboracle|b4d_oracle)\b|(^|\s)(from|import)\s+oracle\b|[\"'/]oracle/")
    offenders = []
    for root, _dirs, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and f != "Makefile":
                continue
            path = os.path.join(root, f)
            for ln, line in enumerate(open(path, errors="replace"), 1):
                code = line.split("#", 1)[0] if (f.endswith(".py") or f == "Makefile") else line.split("//", 1)[0]
                if pat.search(code):
                    offenders.append("%s:%d: %s" % (os.path.relpath(path, pkg), ln, line.strip()))
    assert not offenders, offenders
    needed = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed
