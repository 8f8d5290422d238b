#!/usr/bin/env python
"""Writes profiles/r2_sass_tma.md: which product kernels contain TMA bulk copies / mbarriers in their sm_100a SASS
(cuobjdump -sass on the in-tree libb200spmv.so), and that none contains tensor-core or TMEM instructions."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "singlespmv_b200", "libb200spmv.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
names = [f.split("\n", 1)[0].strip() for f in funcs]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
fam = collections.OrderedDict()
excerpt = None
for f, d in zip(funcs, dem):
    base = re.sub(r"^void ", "", d).split("<")[0].split("(")[0]
    if not base.startswith("b2::") and not base.startswith("(anonymous"):
        continue                                                   # CUB / thrust internals
    blk, syn = len(re.findall(r"\bUBLKCP", f)), len(re.findall(r"\bSYNCS", f))
    mma = len(re.findall(r"\b(UTC\w*MMA|HMMA|HGMMA|LDTM|STTM)\b", f))
    e = fam.setdefault(base, [0, 0, 0, 0])
    e[0] += 1
    e[1] = max(e[1], blk)
    e[2] = max(e[2], syn)
    e[3] += mma
    if excerpt is None and "chunk_stream_kernel<double, double, double, 8, 256, false>" in d:
        excerpt = (d, f)
out = ["# SASS evidence: TMA bulk copies and mbarriers in the product kernels (sm_100a)\n\n",
       "`python scripts/sass_evidence.py` = `cuobjdump -sass singlespmv_b200/libb200spmv.so`, grouped by kernel template.\n",
       "`UBLKCP` = `cp.async.bulk` (1-D TMA bulk copy), `SYNCS` = mbarrier operations (`ARRIVE.TRANS64` = arrive.expect_tx,\n",
       "`PHASECHK` = try_wait).  Tensor-core / TMEM mnemonics (`UTC*MMA`, `HMMA`, `LDTM`, `STTM`) are absent from every\n",
       "kernel: SpMV is not a dense contraction (BASELINE.json north star).\n\n",
       "| kernel template | instantiations | UBLKCP per kernel | SYNCS per kernel | tensor / TMEM instr. |\n|---|---|---|---|---|\n"]
for k, v in fam.items():
    out.append("| `%s` | %d | %d | %d | %d |\n" % (k, v[0], v[1], v[2], v[3]))
if excerpt:
    d, f = excerpt
    keep = [re.sub(r"/\*[0-9a-f]{4}\*/\s*", "", l).split("/*")[0].rstrip()
            for l in f.split("\n") if re.search(r"UBLKCP|SYNCS|LDG|STG|LDS|BAR\.", l)]
    out.append("\n## `%s` (CRS on configs 1 and 5): memory / barrier instructions in program order\n\n```\n" % d.split("(")[0])
    out += [l + "\n" for l in keep[:70]]
    out.append("```\n")
open(os.path.join(ROOT, "profiles", "r2_sass_tma.md"), "w").write("".join(out))
print("".join(out)[:2500])
