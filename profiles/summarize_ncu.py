"""Turns the ncu artefacts gpurun brought back (gpurun_out/) into the text summaries committed here.

    python profiles/summarize_ncu.py r01

Inputs: gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum --clock-control none)
        gpurun_out/prof_score_<tag>.ncu-rep, gpurun_out/prof_hbm_<tag>.ncu-rep   (ncu --set full)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def raw_csv(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarize_rep(rep, out):
    rows = raw_csv(rep)
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep}); per-launch values\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n== {name}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:86s} {r[i]:>18s} {units[i]}\n")
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
            rd = float(r[ir].replace(",", "")) * mult.get(units[ir], 1.0)
            wr = float(r[iw].replace(",", "")) * mult.get(units[iw], 1.0)
            f.write(f"{'traffic = dram read + write':86s} {(rd + wr) / 1e6:18.3f} Mbyte\n")


def summarize_launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 1 --warmup 1 --no-cpu-baseline\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's kernel_ms_per_step\n")
        f.write(f"total kernel time {tot:.1f} us\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:44s} {a[0]:5d} launches {a[1]:12.1f} us {100 * a[1] / tot:6.2f}%\n")


if __name__ == "__main__":
    import os
    tag = sys.argv[1]                                   # suffix of the gpurun_out/ artefacts
    out = sys.argv[2] if len(sys.argv) > 2 else tag     # prefix of the committed summaries
    summarize_launches(f"gpurun_out/launches_{tag}.csv", f"profiles/{out}_launches_summary.txt")
    summarize_rep(f"gpurun_out/prof_score_{tag}.ncu-rep", f"profiles/{out}_score_kernel_ncu.txt")
    summarize_rep(f"gpurun_out/prof_hbm_{tag}.ncu-rep", f"profiles/{out}_hbm_kernels_ncu.txt")
    if os.path.exists(f"gpurun_out/prof_reab_{tag}.ncu-rep"):
        summarize_rep(f"gpurun_out/prof_reab_{tag}.ncu-rep", f"profiles/{out}_reabsorb_kernels_ncu.txt")
