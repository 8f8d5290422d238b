#!/usr/bin/env python
"""Tabulate report blocks, like the reference's log/format.cpp:10-50: each block between the `++++` and `----` lines is
a set of `key value` pairs (first two whitespace-separated tokens of a line); blocks are sorted by nNnz then Matrix and
printed as TSV `Matrix Architecture MatrixFormat Performance(GFLOPS) nRow nCol nNnz`.  `--roofline` appends the keys the
B200 driver adds (EffectiveBW(GB/s), RooflinePct(8000GB/s), KernelTime(us)).  `total_gflops()` is log/sum.sh:4-9."""
import sys

BEGIN, END = "+" * 40, "-" * 40
COLUMNS = ["Matrix", "Architecture", "MatrixFormat", "Performance(GFLOPS)", "nRow", "nCol", "nNnz"]
EXTRA = ["EffectiveBW(GB/s)", "RooflinePct(8000GB/s)", "KernelTime(us)"]


def parse(lines):
    data, cur = [], {}
    for line in lines:
        line = line.rstrip("\n")
        if line == BEGIN:
            cur = {}
        elif line == END:
            data.append(cur)
        else:
            tok = line.split()
            if tok:
                cur[tok[0]] = tok[1] if len(tok) > 1 else ""
    return data


def atoi(s):
    """C atoi: leading integer, 0 when there is none (log/format.cpp:44 sorts with it)."""
    n, s = 0, (s or "").strip()
    sign = -1 if s[:1] == "-" else 1
    for ch in s.lstrip("+-"):
        if not ch.isdigit():
            break
        n = n * 10 + int(ch)
    return sign * n


def table(data, roofline=False):
    data = sorted(data, key=lambda d: (atoi(d.get("nNnz")), d.get("Matrix", "")))
    cols = COLUMNS + (EXTRA if roofline else [])
    return ["\t".join(d.get(c, "") for c in cols) for d in data]


def total_gflops(data):
    return sum(float(d["Performance(GFLOPS)"]) for d in data if d.get("Performance(GFLOPS)"))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if len(args) != 1:
        print("Usage: %s [--roofline] <log file>" % sys.argv[0])
        sys.exit(1)
    for row in table(parse(open(args[0])), "--roofline" in sys.argv):
        print(row)


if __name__ == "__main__":
    main()
