#!/usr/bin/env python
"""Turn the raw captures of scripts/gpu_profile_round.sh (gpurun_out/<tag>_*) into the tracked summaries under
profiles/: shape-mapped launch table, per-family step breakdown, DRAM traffic per conv launch (roofline.traffic of
bench.py, stamped with the commit it was captured on), key ncu metrics of every full capture, SASS opcode histogram.

    python scripts/summarize_profiles.py r2
"""
import collections
import csv
import datetime
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def sh(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


def head():
    """the last commit that changed what was measured (kernel sources, host runtime, bench) -- later commits are docs / evidence"""
    return sh(f"git -C {ROOT} log -1 --format=%h -- v-diffusion-torch_b200 bench.py scripts/prof_kernels.py").strip()


def split_launch_csv(path):
    """one metric per row -> {metric: csv text of that metric only} so the single-metric scripts can read it"""
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    out = collections.defaultdict(list)
    for r in rows:
        out[r["Metric Name"]].append(r)
    return rows[0].keys() if rows else [], out


def write_csv(fields, rows, path):
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(fields), quoting=csv.QUOTE_ALL)
        w.writeheader()
        w.writerows(rows)


def key_metrics(rep, title):
    raw = sh(f"ncu -i {rep} --page raw --csv 2>/dev/null")
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return f"## {title}: no data in {os.path.basename(rep)}\n"
    hdr, units = rows[0], rows[1]
    txt = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        txt.append(f"\n## {title}: {name[:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                txt.append(f"{k:80s} {r[i]} {units[i]}")
    return "\n".join(txt) + "\n"


def main(tag):
    os.makedirs(PR, exist_ok=True)
    stamp = {"git_head": head(), "summarised_utc": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ")}
    lp = os.path.join(GO, f"{tag}_launches_1024rows.csv")
    if os.path.exists(lp):
        fields, per = split_launch_csv(lp)
        tpath = os.path.join(PR, f"{tag}_launches_1024rows.csv")
        write_csv(fields, per["gpu__time_duration.sum"], tpath)
        bpath = os.path.join(GO, f"{tag}_dram_bytes.csv")
        write_csv(fields, per["dram__bytes_read.sum"] + per["dram__bytes_write.sum"], bpath)
        table = sh(f"python {ROOT}/scripts/shape_table.py {tpath} 1")
        brk = sh(f"python {ROOT}/scripts/step_breakdown.py {tpath}")
        open(os.path.join(PR, f"{tag}_shape_table_1024rows.txt"), "w").write(
            f"# commit {stamp['git_head']}; ncu --metrics gpu__time_duration.sum --clock-control none, one denoising step over one\n"
            f"# 1024-row chunk (CIFAR-10 cond, CFG); launches are cold-cache and serialised: compare shares, not absolutes\n" + table + "\n" + brk)
        out = os.path.join(PR, f"{tag}_conv_dram_traffic.json")
        sh(f"python {ROOT}/scripts/conv_traffic.py {bpath} 1024 {out}")
        if os.path.exists(out):
            d = json.load(open(out))
            d.update(stamp)
            json.dump(d, open(out, "w"), indent=1)
        print(table[-400:], brk)
    txt = [f"# commit {stamp['git_head']}; ncu --set full --clock-control none --import-source on, one launch each (scripts/gpu_profile_round.sh)\n"]
    for rep, title in ((f"{tag}_kernels", "dominant kernels at BASELINE configs[1] shapes (1024-row chunk)"),
                       (f"{tag}_kernels_smallk", "epilogue-bound GEMMs (K = 64 / 256)"),
                       (f"{tag}_sampler_step", "fused sampler update, B = 4096 CFG")):
        rp = os.path.join(GO, rep + ".ncu-rep")
        if os.path.exists(rp):
            txt.append(key_metrics(rp, title))
    open(os.path.join(PR, f"{tag}_ncu_key_metrics.txt"), "w").write("\n".join(txt))
    tp = os.path.join(GO, f"{tag}_kernel_timings.txt")
    if os.path.exists(tp):
        open(os.path.join(PR, f"{tag}_kernel_timings.txt"), "w").write(
            f"# commit {stamp['git_head']}; scripts/prof_kernels.py: CUDA-event timings of single launches (not under a profiler)\n" + open(tp).read())
    rp = os.path.join(GO, f"{tag}_kernels.ncu-rep")
    if os.path.exists(rp):
        # per-instruction warp-stall samples of the attention kernel (source page of the same capture)
        page = sh(f"ncu -i {rp} --page source --csv --kernel-name regex:attention_kernel")
        lines = page.splitlines()
        hdr = next((i for i, l in enumerate(lines) if l.startswith('"Address"')), None)
        if hdr is not None:
            rows = list(csv.DictReader(lines[hdr:]))
            reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
            tot = sum(int(r["# Samples"] or 0) for r in rows)
            dur = sh(f"ncu -i {rp} --page raw --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum "
                     f"--kernel-name regex:attention_kernel").strip().splitlines()[-1]
            with open(os.path.join(PR, f"{tag}_attention_stalls.txt"), "w") as f:
                f.write(f"# commit {stamp['git_head']}; ncu --set full --import-source on, persistent attention_kernel<fp16>, N=1024 d=256, 1024 images\n"
                        f"# raw metrics row (gpu__time_duration ms, ...): {dur[-120:]}\n"
                        f"# warp-stall samples by SASS instruction (total {tot}); columns: samples, share, times executed, instruction, top stall reasons\n")
                for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:40]:
                    n = int(r["# Samples"] or 0)
                    top = sorted(((k[6:], int(r[k] or 0)) for k in reasons), key=lambda kv: -kv[1])[:3]
                    f.write(f"{n:7d} {100.0 * n / max(1, tot):5.1f}% {r['Instructions Executed']:>9s}  {r['Source'].strip()[:80]:80s} {top}\n")
    bp = os.path.join(GO, f"{tag}_backward_launches.csv")
    if os.path.exists(bp):
        fields, per = split_launch_csv(bp)
        num = lambda r: float(r["Metric Value"].replace(",", ""))
        mb = lambda r: num(r) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(r["Metric Unit"], 1e-6)
        t = [(r["Kernel Name"], num(r)) for r in per["gpu__time_duration.sum"]]
        rd = [mb(r) for r in per["dram__bytes_read.sum"]]
        wr = [mb(r) for r in per["dram__bytes_write.sum"]]
        unit = per["gpu__time_duration.sum"][0]["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        with open(os.path.join(PR, f"{tag}_backward_launches.txt"), "w") as f:
            f.write(f"# commit {stamp['git_head']}; ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n"
                    "# python scripts/prof_kernels.py --backward --once: GroupNorm backward C256 @32 B=128 (five launches), then conv backward\n"
                    "# (dgrad on conv_gemm_kernel incl. its weight-pack kernels, wgrad + split reduction + bias grad) for 256->256 k3 @32, @16 and 512->256 k3 @32, B=128\n"
                    "# kernel, duration us, DRAM MB read + written\n")
            for (name, dur), r, w in zip(t, rd, wr):
                if name.startswith("void at::") or name.startswith("at::"):      # torch's own input-generation kernels
                    continue
                f.write(f"{name[:70]:70s} {dur * scale:10.1f} {r + w:10.1f}\n")
    # SASS opcode histogram of the shipped library: proves tcgen05 / TMEM / TMA (B200_PROFILING.md)
    sass = sh(f"cuobjdump -sass {ROOT}/v-diffusion-torch_b200/lib/libvdt_b200.so")
    ops = collections.Counter()
    per_fn = collections.defaultdict(collections.Counter)
    fn = "?"
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            dem = sh(f"c++filt {m.group(1)}").strip()
            km = re.search(r"(\w+_kernel(?:<[^>]*>)?)", dem)
            fn = km.group(1) if km else dem.split("(")[0][-60:]
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            ops[op.split(".")[0]] += 1
            if re.match(r"UTC|LDTM|STTM|UTMA|UBLKCP|HMMA|SYNCS|MUFU", op):
                per_fn[fn][op] += 1
    with open(os.path.join(PR, f"{tag}_sass_opcodes.txt"), "w") as f:
        f.write(f"# commit {stamp['git_head']}; cuobjdump -sass v-diffusion-torch_b200/lib/libvdt_b200.so\n")
        f.write("# Blackwell-native evidence: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor loads; no HMMA (mma.sync)\n")
        for fnname, c in sorted(per_fn.items()):
            f.write(f"\n{fnname}\n")
            for op, n in sorted(c.items(), key=lambda kv: -kv[1]):
                f.write(f"    {op:40s} {n}\n")
        f.write("\n# whole library, by base mnemonic\n")
        for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:60]:
            f.write(f"{op:20s} {n}\n")
    print("wrote profiles/" + tag + "_*")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
