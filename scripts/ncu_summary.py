#!/usr/bin/env python
"""Text summary of an ncu report (the metrics DESIGN.md / profiles/README.md quote): python scripts/ncu_summary.py x.ncu-rep"""
import csv
import subprocess
import sys

KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct lts__t_bytes.sum launch__grid_size launch__block_size
launch__registers_per_thread launch__shared_mem_per_block_dynamic sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum sm__cycles_elapsed.max
smsp__average_warp_latency_per_inst_issued.ratio smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum sm__inst_executed_pipe_lsu.sum sm__inst_executed_pipe_alu.sum sm__inst_executed_pipe_fma.sum
sm__inst_executed_pipe_xu.sum sm__throughput.avg.pct_of_peak_sustained_elapsed""".split()


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("## " + r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                if r[i] not in ("", "n/a") and "nan" not in r[i]:
                    print("%-90s %-16s %s" % (k, units[i], r[i]))
        print()


if __name__ == "__main__":
    main()
