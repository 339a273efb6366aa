#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.md
    python profiles/summarize.py full gpurun_out/prof_synth_r01.ncu-rep > profiles/r01_k_synth_full.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__average_warp_latency_issue_stalled_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__shared_mem_per_block_dynamic"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ik].split("(")[0] + " grid=" + r[ig] + " block=" + r[ib]
        agg.setdefault(k, []).append(float(r[iv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | total ms | share | mean us |")
    print("|---|---:|---:|---:|---:|")
    for k, v in agg.items():
        print("| %s | %d | %.3f | %.1f %% | %.1f |" % (k, len(v), sum(v) / 1e6, 100 * sum(v) / tot, sum(v) / len(v) / 1e3))
    print("\n(total %.3f ms over %d launches; ncu per-launch times are cold-cache and serialised: compare shares)" % (tot / 1e6, len(rows) - 1))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("## %s  grid %s block %s\n" % (d["Kernel Name"][1].split("(")[0], d.get("Grid Size", ("", ""))[1], d.get("Block Size", ("", ""))[1]))
        print("| metric | unit | value |\n|---|---|---:|")
        for k in KEYS:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
