#!/usr/bin/env python3
"""Summarise an ncu capture into profiles/: launch list shares + key metrics + SASS opcode mix.

usage: tools/ncu_summary.py <launches.csv|-> <full.ncu-rep> <tag> [summary-name]
Writes profiles/<tag>_launches.txt (skipped with "-"), profiles/<tag>_metrics.txt,
profiles/<tag>_sass_mix.txt and profiles/<summary-name>_ncu_summary.json (default "validity":
profiles/validity_ncu_summary.json is read by bench.py for roofline.traffic).
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PROF = ROOT / "profiles"

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "lts__t_bytes.sum",
]


def main():
    launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    summary_name = sys.argv[4] if len(sys.argv) > 4 else "validity"
    PROF.mkdir(exist_ok=True)
    if launches != "-":
        write_launch_list(launches, tag)
    write_metrics(rep, tag, summary_name)


def write_launch_list(launches, tag):
    rows = list(csv.reader(open(launches)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    per = defaultdict(list)
    for r in rows[hdr + 2:]:
        if len(r) > vi:
            per[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in per.values())
    with open(PROF / f"{tag}_launches.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n# source: {launches}\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{100 * sum(v) / tot:6.2f}%  n={len(v):3d}  avg={sum(v) / len(v) / 1e3:10.2f} us  max={max(v) / 1e3:10.2f} us  {k[:110]}\n")
    print(open(PROF / f"{tag}_launches.txt").read())


def write_metrics(rep, tag, summary_name):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h = rr[0]
    out = {}
    with open(PROF / f"{tag}_metrics.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none, first captured launch of {rr[2][h.index('Kernel Name')][:100]}\n")
        for m in METRICS:
            if m in h:
                val = rr[2][h.index(m)]
                unit = rr[1][h.index(m)]
                f.write(f"{m:90s} {val} {unit}\n")
                out[m] = (val, unit)
    def tobytes(key):
        v, u = out[key]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    summary = {"kernel": rr[2][h.index("Kernel Name")], "duration_us_under_ncu": out["gpu__time_duration.sum"][0],
               "dram_bytes_per_launch": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"),
               "registers": out["launch__registers_per_thread"][0], "source": Path(rep).name}
    (PROF / f"{summary_name}_ncu_summary.json").write_text(json.dumps(summary, indent=1) + "\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sr = list(csv.reader(io.StringIO(src)))
    h = sr[1]
    si, ei, st = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    ops, stalls, tot = Counter(), Counter(), 0
    for r in sr[2:]:
        if len(r) <= ei or r[0] in ("Kernel Name", "Address"):
            if r and r[0] == "Kernel Name":
                break
            continue
        try:
            e = int(r[ei])
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si])
        op = m.group(2) if m else "?"
        ops[op] += e
        stalls[op] += int(r[st] or 0)
        tot += e
    grid = float(out["launch__grid_size"][0].replace(",", "")) if "launch__grid_size" in out else 0
    with open(PROF / f"{tag}_sass_mix.txt", "w") as f:
        f.write(f"# executed warp instructions by opcode (ncu source page), total {tot}\n")
        for op, c in ops.most_common(40):
            f.write(f"{op:10s} {c:12d} {100 * c / tot:6.2f}%  stall-samples {stalls[op]}\n")
    print(open(PROF / f"{tag}_metrics.txt").read())
    print(open(PROF / f"{tag}_sass_mix.txt").read()[:1500])


if __name__ == "__main__":
    main()
