#!/usr/bin/env python
"""profiles/r2_all_formats.md from one default `bench.py` line (every config x format of the same run).
Usage: python scripts/bench_table.py gpurun_out/<default bench>.json [reference-arm.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(sys.argv[1]))
ref = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
out = ["# Round 2: every config x format, one B200, ONE `python bench.py` run (the driver's own command)\n\n",
       "CUDA-event timings, device-resident vectors; frac = compulsory GB/s over the measured copy peak %.1f GB/s (%s); of8000 = over the\n" % (d["roofline"]["peak"], d["roofline"]["peak_source"]),
       "nominal 8 TB/s.  parity = GPU y against the reference's CRS result (oracle/_ref) on the first rows of the same matrix: `exact` = bit-identical,\n",
       "otherwise the largest error relative to sum |a x| (bar: 1e-12 fp64, 1e-5 fp32).  traffic = ncu DRAM bytes of one multiply / compulsory bytes.\n",
       "Clocks during the headline region: %s.\n\n" % json.dumps(d.get("clocks")),
       "| config | format | ms/multiply | GFLOP/s | alg. GB/s | frac | of8000 | traffic ratio | parity | extra |\n|---|---|---|---|---|---|---|---|---|---|\n"]
for key in sorted(d.get("configs", {})):
    c = d["configs"][key]
    if c.get("error"):
        out.append("| %s | – | – | – | – | – | – | – | – | %s |\n" % (key, c["error"]))
    for f, e in c["formats"].items():
        if "error" in e:
            out.append("| %s | %s | – | – | – | – | – | – | – | %s |\n" % (key, f, e["error"][:80]))
            continue
        p = e.get("parity") or {}
        par = "exact" if p.get("bit_identical") else ("%.1e" % p["max_rel_to_mag"] if p else "–")
        extra = []
        if "warm_l2_ms_per_step" in e:
            extra.append("warm L2 %.1f us" % (e["warm_l2_ms_per_step"] * 1e3))
        if "graph_ms_per_step" in e:
            extra.append("in a CUDA graph %.1f us" % (e["graph_ms_per_step"] * 1e3))
        if e.get("options"):
            extra.append(json.dumps(e["options"]))
        tr = "%.2f" % (e["traffic"] / e["alg_bytes"]) if e.get("traffic") else "–"
        out.append("| %s | %s | %.4f | %.1f | %.0f | %.3f | %.3f | %s | %s | %s |\n" % (
            key, f, e["ms_per_step"], e["gflops"], e["alg_gbs"], e["frac"], e["frac_of_8000"], tr, par, "; ".join(extra)))
    if c.get("cusparse"):
        for a, r in c["cusparse"].items():
            if "gflops" in r:
                out.append("| %s | cuSPARSE %s (comparison) | %.4f | %.1f | – | – | – | – | – | cusparseSpMV on the same device arrays |\n" % (key, a, r["ms_per_step"], r["gflops"]))
out.append("\nHeadline (c5 / CRS): %.1f GFLOP/s, %.4f ms, frac %.3f; e2e with host vectors %.1f GFLOP/s (%.2f ms per SpMV, %d + %d bytes over PCIe).\n"
           % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"]))
if "cpu_baseline" in d:
    cb = d["cpu_baseline"]
    out.append("cpu_baseline: %.2f GFLOP/s on %d cores (%s; %s).\n" % (cb["value"], cb["cores"], cb["kind"], cb["sample"]))
if ref:
    out.append("`bench.py --impl reference`: %.2f GFLOP/s (%s).\n" % (ref["value"], ref["config"]["sample"]))
if "gather_ceiling" in d:
    out.append("\nGather ceiling (random 8-byte loads only): %s\n" % json.dumps(d["gather_ceiling"]))
open(os.path.join(ROOT, "profiles", "r2_all_formats.md"), "w").write("".join(out))
print("".join(out))
