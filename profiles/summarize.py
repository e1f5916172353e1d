#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py rep  gpurun_out/prof_k1.ncu-rep  profiles/r1b_k1_full.txt
    python profiles/summarize.py list gpurun_out/launches.csv     profiles/r1b_launches.txt

`rep`  : one block of key counters per profiled launch (from `ncu --set full`).
`list` : per-kernel launch count, total / mean device time and SHARE of the profiled region
         (from `ncu --metrics gpu__time_duration.sum --csv`).
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def ncu_csv(rep, page="raw"):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def do_rep(rep, dst):
    rows = ncu_csv(rep)
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {rep}: key counters per profiled launch (ncu --set full --clock-control none)", ""]
    for r in rows[2:]:
        lines.append(f"kernel: {r[col['Kernel Name']]}   grid {r[col.get('Grid Size', 0)]}  block {r[col.get('Block Size', 0)]}")
        for k in KEYS:
            if k in col and r[col[k]] != "":
                lines.append(f"  {k:88s} {r[col[k]]:>16s} {units[col[k]]}")
        if "dram__bytes_read.sum" in col:
            def mb(k):
                v, u = float(r[col[k]].replace(",", "")), units[col[k]].lower()
                return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
            lines.append(f"  {'traffic = dram read + write (MB per launch)':88s} {mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum'):16.2f} MB")
        lines.append("")
    open(dst, "w").write("\n".join(lines))
    print(f"wrote {dst} ({len(rows) - 2} launches)")


def do_list(path, dst):
    txt = open(path).read().splitlines()
    start = next(i for i, l in enumerate(txt) if l.startswith('"ID"'))
    rows = list(csv.DictReader(txt[start:]))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {path}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:",
             "# compare SHARES, not absolutes)", "",
             f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'mean us':>10s} {'share':>7s}"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:70]:70s} {n:8d} {us:12.1f} {us / n:10.2f} {100 * us / tot:6.1f}%")
    lines.append(f"{'TOTAL':70s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print(f"wrote {dst}")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    {"rep": do_rep, "list": do_list}[mode](src, dst)
