#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an `ncu --set full --import-source on` report.

ncu's CSV source page is SASS-level only; this joins it, instruction by instruction, with `nvdisasm -g` line annotations
of the same cubin (the kernel must come from the library as built when the report was taken).

    python tools/ncu_lines.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING [top_n]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = {n: i for i, n in enumerate(rows[hdr])}
    inst = []
    for r in rows[hdr + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):     # the next kernel instance of the report
            break
        if len(r) > 5:
            inst.append((r[H["Source"]].strip(), int(r[H["Instructions Executed"]] or 0), int(r[H["# Samples"]] or 0)))
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
    # locate the kernel's section
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
    lines, cur = [], None
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith("\t.section"):
            if lines:
                break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        elif re.search(r"/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    n = min(len(lines), len(inst))
    if len(lines) != len(inst):
        print(f"warning: {len(lines)} instructions in the cubin, {len(inst)} in the report", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0])
    for k in range(n):
        agg[lines[k]][0] += inst[k][1]
        agg[lines[k]][1] += inst[k][2]
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"{'line':>24} {'inst %':>7} {'samples %':>9}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        name = f"{key[0]}:{key[1]}" if key else "?"
        print(f"{name:>24} {100 * v[0] / ti:7.2f} {100 * v[1] / ts:9.2f}")


if __name__ == "__main__":
    main()
