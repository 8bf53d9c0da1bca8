#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_x.ncu-rep [more.ncu-rep] > profiles/rNN_x.md
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def launches(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 10]
    hdr = rows[0]
    agg = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        name = d["Kernel Name"]
        if len(name) > 100:
            name = name[:97] + "..."
        key = (name, d["Block Size"], d["Grid Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| launches | total ms | share | block | grid | kernel |")
    print("|---:|---:|---:|---|---|---|")
    for (name, blk, grd), (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {cnt} | {ns / 1e6:.3f} | {100 * ns / tot:.1f} % | {blk} | {grd} | `{name}` |")
    print(f"\ntotal device time of the listed launches: {tot / 1e6:.3f} ms "
          "(ncu serialises launches and runs them cold-cache: compare shares, not absolutes)")


def full(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            print(f"### `{d['Kernel Name'][0]}`  ({p.split('/')[-1]})\n")
            print("| metric | value | unit |")
            print("|---|---:|---|")
            for k in KEEP:
                if k in d:
                    print(f"| {k} | {d[k][0]} | {d[k][1]} |")
            print("\nwarp stall reasons (warps per issue-active cycle):\n")
            print("| reason | value |")
            print("|---|---:|")
            st = [(h[len(STALL):].replace("_per_issue_active.ratio", ""), float(v[0])) for h, v in d.items()
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
            for name, v in sorted(st, key=lambda t: -t[1]):
                if v >= 0.01:
                    print(f"| {name} | {v:.3f} |")
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
