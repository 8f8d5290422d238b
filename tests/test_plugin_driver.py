"""The C++ host side: singlespmv_b200/plugin/ mirrors the reference's plugin interface (OptimizeProblem /
SpMV, src/opt_crs.h:15-18) and driver loop (src/main.cpp:17-209) on top of the C-ABI.  One binary per
format, like the reference.  CPU: they build, link against libb200spmv.so and refuse to run without a
device.  GPU: they run the reference's own fixtures (rebuilt as .mtx from the committed goldens), pass
the reference's in-binary verification twice and print the reference's report block."""
import os
import subprocess

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "singlespmv_b200", "plugin", "bin")
FORMATS = ["crs", "coo", "ell", "jds", "dia", "ss", "css", "csr5", "hyb"]


@pytest.fixture(scope="module")
def built():
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "singlespmv_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "singlespmv_b200", "plugin")])
    return BIN


def write_mtx(path, g, shuffle_seed=None):
    """Matrix-Market text the way the reference's loader reads it (src/util.cpp:36-50): comment lines, then
    'M N L', then L 1-based triples -- in shuffled order, the loader sorts."""
    row, col, val = g["in_row"], g["in_col"], g["in_val"]
    order = np.arange(len(row))
    if shuffle_seed is not None:
        np.random.default_rng(shuffle_seed).shuffle(order)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% written by tests/test_plugin_driver.py\n")
        f.write("%d %d %d\n" % (int(g["nRow"]), int(g["nCol"]), len(row)))
        for i in order:
            f.write("%d %d %r\n" % (row[i] + 1, col[i] + 1, float(val[i])))


def parse_report(out):
    lines = out.splitlines()
    a, b = lines.index("+" * 40), lines.index("-" * 40)
    kv = {}
    for ln in lines[a + 1:b]:
        k, v = ln.strip().split("\t", 1)
        kv[k.strip()] = v.strip()
    return kv


def test_driver_binaries_build(built):
    for f in FORMATS:
        for suffix in ("", "_dev"):
            exe = os.path.join(built, "spmv_b200_%s%s" % (f, suffix))
            assert os.access(exe, os.X_OK), exe
            r = subprocess.run([exe], capture_output=True, text=True)
            assert r.returncode == 1 and "Usage" in r.stdout          # src/main.cpp:19-22


def test_driver_refuses_without_gpu(built, tmp_path):
    import singlespmv_b200 as sp
    if sp.device_count() > 0:
        pytest.skip("a CUDA device is present")
    mtx = str(tmp_path / "m.mtx")
    write_mtx(mtx, load_golden("fixture_3x3"))
    r = subprocess.run([os.path.join(built, "spmv_b200_crs"), mtx], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FORMATS)
def test_driver_runs_reference_fixtures(built, tmp_path, fmt):
    env = dict(os.environ, SPMV_MIN_SECONDS="0.02", SPMV_NTRY="2")
    for name in ("fixture_10x10", "mini_rmat_s9"):
        g = load_golden(name)
        mtx = str(tmp_path / (name + ".mtx"))
        write_mtx(mtx, g, shuffle_seed=1)
        for suffix in ("", "_dev"):
            exe = os.path.join(built, "spmv_b200_%s%s" % (fmt, suffix))
            r = subprocess.run([exe, mtx], capture_output=True, text=True, timeout=120, env=env)
            assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-500:]
            assert "invalid result" not in r.stdout
            assert r.stderr.count("Verifying") == 2                   # src/main.cpp:40-56
            kv = parse_report(r.stdout)
            assert kv["Architecture"] == "GPU" and kv["MatrixFormat"] == fmt.upper()
            assert kv["Matrix"] == name + ".mtx"
            assert (int(kv["nRow"]), int(kv["nCol"]), int(kv["nNnz"])) == (int(g["nRow"]), int(g["nCol"]), len(g["in_row"]))
            assert float(kv["Performance(GFLOPS)"]) > 0 and int(kv["AlgBytes"]) > 0


@pytest.mark.gpu
def test_driver_synth_shape(built):
    exe = os.path.join(built, "spmv_b200_dia_dev")
    env = dict(os.environ, SPMV_MIN_SECONDS="0.02", SPMV_NTRY="2")
    r = subprocess.run([exe, "synth:box3d27:24"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    kv = parse_report(r.stdout)
    assert int(kv["nRow"]) == 24 ** 3 and int(kv["nNnz"]) == (3 * 24 - 2) ** 3 and kv["VectorResidency"] == "device"


@pytest.mark.gpu
def test_driver_report_keys_of_the_reference(built):
    """Keys the reference's report carries and log/format consumers read (src/main.cpp:155-174,200-206): nThread for
    every format; nStep + StepCount-xx for SS; MulPerf / SumPerf for the -DPROFILING builds of SS and CSS."""
    env = dict(os.environ, SPMV_MIN_SECONDS="0.02", SPMV_NTRY="2")
    # 27 entries per row, W = 4: rows span chains of whole segments -> at least two fold steps (src/opt_ss.cpp:91-142)
    r = subprocess.run([os.path.join(built, "spmv_b200_ss_dev"), "synth:box3d27:12"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    kv = parse_report(r.stdout)
    assert kv["nThread"] == "1" and kv["nGPU"] == "1"
    n = int(kv["nStep"])
    assert n >= 1 and all(("StepCount-%02d" % i) in kv for i in range(n))
    assert sum(int(kv["StepCount-%02d" % i]) for i in range(n)) > 0
    for exe, keys in (("spmv_b200_ss_prof", ("MulPerf", "SumPerf")), ("spmv_b200_css_prof", ("MulPerf(GFLOPS)", "SumPerf(GFLOPS)"))):
        r = subprocess.run([os.path.join(built, exe), "synth:box3d27:12"], capture_output=True, text=True, timeout=120, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        kv = parse_report(r.stdout)
        assert all(float(kv[k]) > 0 for k in keys), kv
