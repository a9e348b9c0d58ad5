#!/usr/bin/env python3
"""Turns an `ncu --set full` report of K1 into the text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_XXX.ncu-rep > profiles/r01_k1_XXX.txt
"""
import collections
import csv
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main(rep):
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
    print(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]}")
    print(f"kernel: {d.get('Kernel Name', ('?',))[0]}  grid {d.get('Grid Size', ('?',))[0]} block {d.get('Block Size', ('?',))[0]}")
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    for k in keys:
        if k in d:
            print(f"{k:75s} {d[k][0]:>16s} {d[k][1]}")
    try:
        rd = float(d["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]
        print(f"{'traffic = dram read + write bytes per launch':75s} {rd + wr:16.0f} byte")
    except Exception:
        pass
    st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(v) for h, v in zip(hdr, vals)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(st.values()) or 1
    print("\n# warp stall sampling (all samples)")
    for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
        print(f"  {k:24s} {v:9d} {100 * v / tot:5.1f}%")

    src = page(rep, "source")
    h2, data = src[1], src[2:]
    ia, isamp = h2.index("Instructions Executed"), h2.index("# Samples")
    ex = [int(r[ia]) for r in data]
    sm = [int(r[isamp]) for r in data]
    te, ts = sum(ex) or 1, sum(sm) or 1
    print(f"\n# SASS: {len(data)} instructions, {te:.4e} warp-instructions executed")
    mix = collections.Counter()
    for r, e in zip(data, ex):
        t = r[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        mix[op] += e
    print("opcode mix (executed): " + ", ".join(f"{op} {100 * v / te:.1f}%" for op, v in mix.most_common(14)))
    print("\n# hottest regions (consecutive SASS lines with equal execution counts, >= 1.5% of instructions)")
    print("  lines      n   exec/line   %inst  %stall-samples  first instruction")
    start = 0
    for i in range(1, len(ex) + 1):
        if i == len(ex) or abs(ex[i] - ex[start]) > 0.12 * max(ex[i], ex[start], 1):
            e, s = sum(ex[start:i]), sum(sm[start:i])
            if e / te > 0.015:
                print(f"  {start:4d}-{i - 1:4d} {i - start:4d}  {ex[start]:.3e}  {100 * e / te:5.1f}%  {100 * s / ts:5.1f}%          {data[start][1].strip()[:60]}")
            start = i


if __name__ == "__main__":
    main(sys.argv[1])
