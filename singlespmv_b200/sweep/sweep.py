#!/usr/bin/env python
"""Experiment sweep for the `gpu` architecture, in the reference's own workflow.

The reference drives its research sweeps from script/todo.csv -- rows `arch,prefix,option`, one binary per row built
with `make <arch> PREFIX=<prefix> OPTION=<option>` and run over every matrix, output appended to log/<arch>-<prefix>.tsv,
a `<logfile>.lock` symlink making re-runs skip finished rows (script/eval_sbatch.sh:1-59) -- and tabulates the
`++++ / ----` report blocks with log/format.cpp.  This does the same for the B200 plugin, without Slurm:

    python -m singlespmv_b200.sweep.sweep todo.csv --matrices synth:lap2d5:1024 path/to/*.mtx --log-dir log
    python -m singlespmv_b200.sweep.log_format log/gpu-b200-css-w4-n4.tsv

A row looks like   gpu,b200-css-w4-n4,-DOPT_B200 -DB200_FORMAT=CSS -DSEGMENT_WIDTH=4 -DN_BLOCK=4 -DB200_DEVICE_RESIDENT
`gen_todo()` enumerates format x tunables the way script/gen_todo.sh:8-47 enumerates SEGMENT_WIDTH and N_BLOCK.
"""
import argparse
import os
import shlex
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
PLUGIN = os.path.join(PKG, "plugin")


def read_todo(path):
    """script/eval_sbatch.sh:2 -- drop comment and blank lines; split on the first two commas only."""
    rows = []
    for line in open(path):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        arch, prefix, option = line.split(",", 2)
        rows.append((arch.strip(), prefix.strip(), option.strip()))
    return rows


def gen_todo(formats=("CRS", "COO", "ELL", "JDS", "DIA", "CSR5"), widths=(1, 4, 32, 1024), n_blocks=(1, 2, 3, 4, 8, 16),
             sigmas=(0, 4, 8, 16, 32)):
    """The experiment matrix: every format, SS over SEGMENT_WIDTH, CSS over SEGMENT_WIDTH x N_BLOCK, CSR5 over sigma."""
    base = "-DOPT_B200 -DB200_DEVICE_RESIDENT"
    rows = [("gpu", "b200-%s" % f.lower(), "%s -DB200_FORMAT=%s" % (base, f)) for f in formats if f != "CSR5"]
    rows += [("gpu", "b200-ss-w%d" % w, "%s -DB200_FORMAT=SS -DSEGMENT_WIDTH=%d" % (base, w)) for w in widths]
    rows += [("gpu", "b200-css-w%d-n%d" % (w, n), "%s -DB200_FORMAT=CSS -DSEGMENT_WIDTH=%d -DN_BLOCK=%d" % (base, w, n))
             for w in widths[:2] for n in n_blocks]
    rows += [("gpu", "b200-csr5-s%d" % s, "%s -DB200_FORMAT=CSR5 -DB200_SIGMA=%d" % (base, s)) for s in sigmas]
    return rows


def build(prefix, option, bin_dir, cuda="/usr/local/cuda"):
    """One binary per row, like `make <arch> PREFIX=... OPTION=...` (reference Makefile:10-21)."""
    os.makedirs(bin_dir, exist_ok=True)
    exe = os.path.join(bin_dir, "%s-spmv.gpu" % prefix)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++11", "-O2", "-DGPU", "-I" + PLUGIN, "-I" + os.path.join(PLUGIN, "standalone"),
           "-I" + os.path.join(ROOT, "include"), "-I" + cuda + "/include"]
    cmd += shlex.split(option)
    cmd += [os.path.join(PLUGIN, f) for f in ("main_b200.cpp", "opt_b200.cpp", os.path.join("standalone", "util.cpp"))]
    cmd += ["-o", exe, "-L" + PKG, "-lb200spmv", "-L" + cuda + "/lib64", "-lcudart", "-Wl,-rpath," + PKG,
            "-Wl,-rpath," + cuda + "/lib64"]
    subprocess.check_call(cmd)
    return exe


def run(rows, matrices, log_dir, bin_dir, force=False, env=None):
    os.makedirs(log_dir, exist_ok=True)
    done = []
    for arch, prefix, option in rows:
        if arch != "gpu":
            print("Invalid arch : %s" % arch, file=sys.stderr)        # script/eval_sbatch.sh:55-57
            continue
        logfile = os.path.join(log_dir, "%s-%s.tsv" % (arch, prefix))
        lock = logfile + ".lock"
        if os.path.lexists(lock) and not force:                        # script/eval_sbatch.sh:14-18
            print("Skip : %s" % logfile)
            continue
        if not os.path.lexists(lock):
            os.symlink(os.path.basename(logfile), lock)
        exe = build(prefix, option, bin_dir)
        with open(logfile, "a") as log:
            for m in matrices:
                r = subprocess.run([exe, m], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
                log.write(r.stdout)
                if r.returncode != 0:
                    open(logfile + ".err", "a").write("%s: rc=%d\n%s\n" % (m, r.returncode, r.stderr[-2000:]))
        done.append(logfile)
    return done


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("todo", nargs="?", help="todo.csv (arch,prefix,option); omit with --gen")
    ap.add_argument("--gen", action="store_true", help="print the generated experiment matrix and exit")
    ap.add_argument("--matrices", nargs="+", default=[], help=".mtx paths and/or synth:kind:p0[:p1] specs")
    ap.add_argument("--log-dir", default="log")
    ap.add_argument("--bin-dir", default=os.path.join(PLUGIN, "bin"))
    ap.add_argument("--force", action="store_true", help="ignore .lock files")
    a = ap.parse_args()
    if a.gen:
        for row in gen_todo():
            print(",".join(row))
        return
    for f in run(read_todo(a.todo), a.matrices, a.log_dir, a.bin_dir, a.force):
        print(f)


if __name__ == "__main__":
    main()
