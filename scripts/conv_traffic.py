#!/usr/bin/env python
"""Reduce an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` CSV of one denoising step to the average
DRAM bytes per conv_gemm launch (roofline.traffic in bench.py).  usage: conv_traffic.py launches.csv rows out.json"""
import csv
import json
import sys


def main(path, rows, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = {}
    order = []
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r.get("Metric Unit", "byte"), 1.0)
        key = r["ID"]
        if key not in per:
            per[key] = [r["Kernel Name"], 0.0]
            order.append(key)
        per[key][1] += v
    names = [per[k] for k in order]
    marks = [i for i, (n, _) in enumerate(names) if "sampler_begin_step" in n]
    step = names[marks[0]:marks[1]]
    conv = [b for n, b in step if "conv_gemm" in n]
    gn = [b for n, b in step if "groupnorm_kernel" in n]
    at = [b for n, b in step if "attention" in n]
    res = {"rows": int(rows), "conv_launches": len(conv), "avg_dram_bytes_per_conv_launch_scaled_to_rows": sum(conv) / len(conv),
           "total_dram_bytes_conv": sum(conv), "total_dram_bytes_groupnorm": sum(gn), "total_dram_bytes_attention": sum(at),
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one denoising step at `rows` UNet rows"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
