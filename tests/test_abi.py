"""CPU checks of the drop-in boundary: libb200spmv.so loads, exports every symbol include/b200spmv.h
declares, validates arguments, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    out = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            out += re.findall(r"B200SPMV_API[^;(]*?\b(b200spmv_\w+)\s*\(", src)
    return out


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("b200spmv_create", "b200spmv_convert_coo_host", "b200spmv_convert_coo_device",
                 "b200spmv_multiply", "b200spmv_multiply_host", "b200spmv_get_array", "b200spmv_destroy"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import singlespmv_b200 as sp
    lib = C.CDLL(sp.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.b200spmv_version() == 100


def test_no_torch_types_in_signatures():
    src = open(os.path.join(ROOT, "include", "b200spmv.h")).read()
    assert "torch" not in src.lower() and "at::" not in src


def test_argument_validation_without_gpu():
    import singlespmv_b200 as sp
    with pytest.raises(sp.B200SpmvError):
        sp.SpMatOpt("ss", segment_width=3)            # W must be a power of two (opt_ss.cpp:272)
    with pytest.raises(KeyError):
        sp.SpMatOpt("nonsense")
    m = sp.SpMatOpt("crs")
    with pytest.raises(sp.B200SpmvError) as e:        # multiply before convert
        m.multiply(0, 0)
    assert e.value.status == -4 or e.value.status == -1


def test_no_cpu_fallback():
    """Without a device, conversion and multiply must fail loudly -- never compute on the host."""
    import singlespmv_b200 as sp
    if sp.device_count() > 0:
        pytest.skip("a CUDA device is present")
    A = sp.SpMat(3, 3, [0, 1, 2], [0, 1, 2], [1.0, 2.0, 3.0])
    with pytest.raises(sp.B200SpmvError) as e:
        sp.OptimizeProblem(A, sp.Vec(np.ones(3)), "crs")
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(sp.B200SpmvError):
        sp.DeviceCoo("lap2d5", 8)


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under singlespmv_b200/ or include/ may reference it."""
    bad = []
    for base in ("singlespmv_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep) or "__pycache__" in dp:
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")) or fn == "Makefile":
                    txt = open(os.path.join(dp, fn), errors="replace").read()
                    if re.search(r"liboracle|oracle_lib|orc_|libref_|oracle/_ref", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_reference_vectors_match_the_oracle(oracle):
    import singlespmv_b200 as sp
    x, y = sp.reference_vectors(17, 9, 3)
    xo, yo = oracle.reference_vectors(17, 9, 3)
    assert np.array_equal(x, xo) and np.array_equal(y, yo)
    assert x[0] == 0.56138017520372763
