#!/usr/bin/env python
"""Summarise an .ncu-rep (one captured kernel) into the handful of numbers DESIGN.md / profiles/ quote.
   python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [samples_per_launch] > profiles/rNN_X_ncu.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        print(f"# kernel: {rec.get('Kernel Name')}  (ncu --set full --clock-control none; one launch, ~40 replays)")
        vals = {}
        for k in KEYS:
            if k in rec:
                u = units[hdr.index(k)]
                vals[k] = rec[k]
                print(f"{k:92s} {rec[k]:>18s} {u}")
        if nsamp:
            try:
                inst = float(vals["smsp__inst_executed.sum"].replace(",", ""))
                print(f"{'derived: warp instructions per 32 samples':92s} {inst / (nsamp / 32):18.1f}")
                cyc = float(vals["sm__cycles_elapsed.avg"].replace(",", ""))
                pct = float(vals["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"])
                # fp64 pipe: 64 lanes/clk/SM = 2 warp-instructions/clk/SM at 100 %
                fp64 = pct / 100 * cyc * 2 * 148
                print(f"{'derived: FP64-pipe warp instructions per 32 samples':92s} {fp64 / (nsamp / 32):18.1f}")
                rd = float(vals["dram__bytes_read.sum"].replace(",", ""))
                unit = units[hdr.index("dram__bytes_read.sum")]
                mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(unit, 1)
                print(f"{'derived: DRAM bytes read per sample':92s} {rd * mult / nsamp:18.2f}")
            except (KeyError, ValueError) as e:
                print("# derived metrics unavailable:", e)


if __name__ == "__main__":
    main()
