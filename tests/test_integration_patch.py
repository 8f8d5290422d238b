"""The drop-in claim of INTEGRATION.md, executed: the reference's OWN driver and registry (src/main.cpp, src/opt.h,
src/opt.cpp, src/util.cpp -- copied to a temp directory, never into the repo) get the three-hunk patch a maintainer would
apply, and are compiled with -DOPT_B200 against singlespmv_b200/plugin/opt_b200.{h,cpp} + libb200spmv.so.  The binary
that comes out is the reference's main() driving the B200 plugin through OptimizeProblem / SpMV.

Needs /root/reference, so it runs in the build container only (no GPU there: the binary is built, linked and asked for
its usage line; running it is what tests/test_plugin_driver.py does with the standalone twin of the driver)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
PLUGIN = os.path.join(ROOT, "singlespmv_b200", "plugin")
PKG = os.path.join(ROOT, "singlespmv_b200")
CUDA = "/usr/local/cuda"

needs_ref = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference sources not mounted")


def patched_tree(tmp):
    """Copy the reference's src/ and apply the patch of INTEGRATION.md section 2."""
    dst = os.path.join(str(tmp), "src")
    shutil.copytree(REF_SRC, dst)
    # --- src/opt.h: one more format branch
    p = os.path.join(dst, "opt.h")
    s = open(p).read()
    anchor = '#ifdef OPT_CUSPARSE\n#include "opt_cusparse.h"\n#endif\n'
    assert anchor in s
    s = s.replace(anchor, anchor + '#ifdef OPT_B200\n#include "opt_b200.h"\n#endif\n')
    open(p, "w").write(s)
    # --- src/opt.cpp: the GPU architecture includes the B200 plugin instead of the cuSPARSE 6.5 one
    p = os.path.join(dst, "opt.cpp")
    s = open(p).read()
    anchor = '#ifdef GPU\n    #include "opt_cusparse.cpp"\n#endif\n'
    assert anchor in s
    s = s.replace(anchor, '#ifdef GPU\n  #ifdef OPT_B200\n    #include "opt_b200.cpp"\n  #else\n    #include "opt_cusparse.cpp"\n  #endif\n#endif\n')
    open(p, "w").write(s)
    # --- src/main.cpp: the MatrixFormat line of the report (README.md:5-8 step 3)
    p = os.path.join(dst, "main.cpp")
    s = open(p).read()
    anchor = '#ifdef OPT_CUSPARSE\n    printf("%25s\\t%s\\n", "MatrixFormat", "CUSPARSE");\n    isDefinedFormat = true;\n#endif\n'
    assert anchor in s
    s = s.replace(anchor, anchor + '#ifdef OPT_B200\n    printf("%25s\\t%s\\n", "MatrixFormat", B200FormatName());\n    isDefinedFormat = true;\n#endif\n')
    open(p, "w").write(s)
    return dst


def build(src, fmt, exe, extra=()):
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(PKG, "csrc")])
    cmd = ["/usr/bin/g++", "-std=c++11", "-O2", "-fopenmp", "-w", "-Drestrict=__restrict__",
           "-DGPU", "-DVERIFY", "-DINDEX_32", "-DALIGNMENT=32", "-DOPT_B200", "-DB200_FORMAT=" + fmt, *extra,
           "-I" + src, "-I" + PLUGIN, "-I" + os.path.join(ROOT, "include"), "-I" + CUDA + "/include",
           os.path.join(src, "main.cpp"), os.path.join(src, "util.cpp"), os.path.join(src, "opt.cpp"),
           "-o", exe, "-L" + PKG, "-lb200spmv", "-L" + CUDA + "/lib64", "-lcudart",
           "-Wl,-rpath," + PKG, "-Wl,-rpath," + CUDA + "/lib64"]
    subprocess.check_call(cmd)
    return exe


@needs_ref
@pytest.mark.parametrize("fmt,extra", [("CRS", ()), ("CSS", ("-DSEGMENT_WIDTH=4", "-DN_BLOCK=2")), ("DIA", ())])
def test_reference_driver_builds_with_the_plugin(tmp_path, fmt, extra):
    src = patched_tree(tmp_path)
    exe = build(src, fmt, str(tmp_path / ("spmv_" + fmt)), extra)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stdout                  # the reference's own usage line, src/main.cpp:19-22
    # the reference's symbols are in the binary next to the plugin's
    syms = subprocess.run(["nm", "-C", exe], capture_output=True, text=True).stdout
    for s in ("OptimizeProblem(SpMat const&, Vec const&, SpMatOpt&, VecOpt&)", "SpMV", "VerifyResult(SpMat const&, Vec const&, Vec const&)",
              "LoadSparseMatrix", "b200spmv_multiply_host"):
        assert s in syms, s
