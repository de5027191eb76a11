#!/usr/bin/env python
"""Sum an ncu launch list (gpu__time_duration) per kernel family for ONE denoising step (the launches between
two sampler_begin_step kernels), printing per-family totals and the largest conv launches."""
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r.get("Metric Unit", "ns"), 1.0)
        seq.append((r["Kernel Name"], v))
    marks = [i for i, (n, _) in enumerate(seq) if "sampler_begin_step" in n]
    if len(marks) < 2:
        print("need two sampler_begin_step launches in the list")
        return
    step = seq[marks[0]:marks[1]]
    fam = {}
    for n, v in step:
        k = ("conv" if "conv_gemm" in n else "gn_finalize" if "finalize" in n else "groupnorm" if "groupnorm" in n
             else "attention" if "attention" in n else "other")
        fam.setdefault(k, [0, 0.0])
        fam[k][0] += 1
        fam[k][1] += v
    tot = sum(v for _, v in fam.values())
    print(f"# {path}: one step = {len(step)} launches, {tot / 1e3:.3f} ms (ncu-serialised)")
    for k, (c, v) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:12s} n={c:4d} total_ms={v / 1e3:8.3f} share={v / tot * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
