#!/usr/bin/env python3
"""Summarise an .ncu-rep (from `ncu --set full --import-source on`) into the text kept under profiles/.

usage: tools/ncu_summary.py report.ncu-rep cells_per_launch [lanes_per_warp_instr_cells]
Prints, per captured launch: duration, DRAM bytes, issue utilisation, pipe utilisation, occupancy, stall reasons,
and the dynamic SASS opcode mix per 32 cell-updates (from the per-instruction execution counts of the source page).
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "sm__cycles_elapsed.avg",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"== {name}")
        vals = {}
        for k in KEYS:
            if k in hdr:
                vals[k] = r[hdr.index(k)]
                print(f"   {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
        stalls = []
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", h)
            if m:
                try:
                    stalls.append((float(r[i]), m.group(1) or m.group(2)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("   stall reasons (warp-cycles per issued instruction): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))
        if cells and "smsp__inst_executed.sum" in vals:
            print(f"   warp-instructions per 32 cell-updates = {float(vals['smsp__inst_executed.sum']) / (cells / 32):.1f}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    for si, s in enumerate(secs[:1]):
        hdr = src[s + 1]
        ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
        isamp = hdr.index("# Samples")
        end = secs[si + 1] if si + 1 < len(secs) else len(src)
        ops, samp = collections.Counter(), collections.Counter()
        for r in src[s + 2:end]:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
            if m and r[ia].isdigit():
                ops[m.group(2)] += int(r[ia])
                samp[m.group(2)] += int(r[isamp])
        tot = sum(ops.values())
        print(f"== dynamic SASS mix of launch 0: {tot} warp-instructions")
        for op, c in ops.most_common(28):
            per = f"{c / (cells / 32):7.1f} per 32 cells" if cells else ""
            print(f"   {op:10s} {100.0 * c / tot:5.1f}%  {per}   stall samples {samp[op]}")


if __name__ == "__main__":
    main()
