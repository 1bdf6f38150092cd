#!/usr/bin/env python3
"""Per-source-line view of an ncu report (needs -lineinfo and --import-source on): executed warp instructions and stall
samples per CUDA source line, top N.  usage: tools/ncu_lines.py <report.ncu-rep> [N]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue
        try:
            smp, ins = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        a = agg.setdefault((cur, int(r[0])), [0, 0, r[1][:110]])
        a[0] += smp
        a[1] += ins
    tot, ti = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
    print(f"# {rep}: {tot} stall samples, {ti} warp instructions attributed (inlined lines are counted at every level)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{a[0] / tot * 100:5.1f}% smp {a[1] / ti * 100:5.1f}% ins  {k[0]}:{k[1]}  {a[2]}")


if __name__ == "__main__":
    main()
