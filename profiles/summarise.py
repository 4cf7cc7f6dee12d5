#!/usr/bin/env python
"""Turn the ncu captures of a round into the text summaries kept in this directory.

    python profiles/summarise.py full   gpurun_out/prof_r01d.ncu-rep  r01d   "<the ncu command line>"
    python profiles/summarise.py launch gpurun_out/launches_r01d.csv  r01d   "<the ncu command line>"

`full`   one profiles/<tag>_ncu_<kernel>.txt per captured kernel (selected raw metrics + the source lines with the
         most stall samples) and profiles/traffic.json (dram read + write bytes per launch, read by bench.py).
`launch` profiles/<tag>_ncu_launches_summary.txt: per-kernel totals of the gpu__time_duration pass.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]
STAGE = {"k_flood2": "s1.flood", "k_rag_accumulate": "s2.rag", "k_agglomerate_par": "s2.agglomerate", "k_tile_front": "s1.tile_front",
         "k_mask_bits_u8": "s1.mask_bits", "k_finalize": "s1.finalize", "k_fragstats": "s1.fragstats", "k_relabel_dense": "s3.relabel"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short(name):
    m = re.search(r"(k_\w+)", name)
    return m.group(1) if m else name.split("(")[0]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, tag, cmd):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    for r in data:
        name = short(r[col["Kernel Name"]])
        lines = [f"# {cmd}", f"{'Kernel Name':90s} {r[col['Kernel Name']]}"]
        for m in METRICS:
            if m in col:
                lines.append(f"{m:90s} {r[col[m]]} {units[col[m]]}")
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_write.sum"]]]
        lines.append(f"{'dram read + write per launch':90s} {rd + wr:.0f} byte")
        if name in STAGE:
            traffic[STAGE[name]] = rd + wr
        path = os.path.join(HERE, f"{tag}_ncu_{name.replace('k_', '')}.txt")
        with open(path, "w") as f:
            f.write("\n".join(lines) + "\n")
        print("wrote", path)
    # hottest source lines per kernel (stall samples aggregated per CUDA source line, inlined headers included)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True,
                         text=True).stdout
    per_kernel = {}
    for b in re.split(r'(?m)^"File Path",', src)[1:]:
        rr = list(csv.reader(io.StringIO('"File Path",' + b)))
        fn = [r for r in rr[:4] if r and r[0] == "Function Name"]
        hd = next((r for r in rr[:6] if r and r[0] == "Line No"), None)
        if not fn or hd is None:
            continue
        i_smp, i_ins = hd.index("# Samples"), hd.index("Instructions Executed")
        fname = os.path.basename(rr[0][1])
        for r in rr:
            if len(r) > i_ins and r[0].isdigit() and r[2] == "-" and r[i_smp].isdigit():
                per_kernel.setdefault(short(fn[0][1]), []).append((int(r[i_smp]), int(r[i_ins] or 0), fname, r[0], r[1].strip()))
    for kname, body in per_kernel.items():
        tot = sum(b[0] for b in body) or 1
        tot_i = sum(b[1] for b in body) or 1
        path = os.path.join(HERE, f"{tag}_ncu_{kname.replace('k_', '')}.txt")
        with open(path, "a") as f:
            f.write("# source lines with the most warp-stall samples: share of samples, share of executed instructions\n")
            for smp, ins, fname, line, text in sorted(body, key=lambda b: -b[0])[:16]:
                f.write(f"{100.0 * smp / tot:6.2f}% {100.0 * ins / tot_i:6.2f}%  {fname}:{line:<5s} {text[:130]}\n")
    if traffic:
        traffic["_source"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture summarised in "
                              f"profiles/{tag}_ncu_*.txt (config 2, 1 B200)")
        with open(os.path.join(HERE, "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
        print("wrote traffic.json", traffic)


def launch(path, tag, cmd):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in rows[1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        u = r[col["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        n = short(r[col["Kernel Name"]])
        if n == "k_synth":      # the synthetic input generator runs before the timed region
            continue
        t = re.search(r"<(.*)>", r[col["Kernel Name"]])
        if t and n in ("k_agglomerate_par", "k_flood2", "k_mask_rowdist", "k_maxfilt_xy_t", "k_rag_accumulate"):
            n += "<" + t.group(1)[:40] + ">"
        a = acc.setdefault(n, [0.0, 0])
        a[0] += ms
        a[1] += 1
    return acc


def launch_main(path, tag, cmd, steps):
    acc = launch(path, tag, cmd)
    tot = sum(a[0] for a in acc.values())
    n = sum(a[1] for a in acc.values())
    out = os.path.join(HERE, f"{tag}_ncu_launches_summary.txt")
    with open(out, "w") as f:
        f.write(f"# {cmd}\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's stage_ms, not absolutes\n")
        f.write(f"# {n} launches over {steps} steps, {tot / steps:.3f} ms per step\n")
        for k, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{a[0] / steps:10.3f} ms {a[1] / steps:6.1f}x {100 * a[0] / tot:5.1f}%  {k}\n")
    print("wrote", out)


if __name__ == "__main__":
    mode, path, tag, cmd = sys.argv[1:5]
    if mode == "full":
        full(path, tag, cmd)
    else:
        launch_main(path, tag, cmd, int(sys.argv[5]) if len(sys.argv) > 5 else 1)
