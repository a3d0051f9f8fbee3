"""Condense an `ncu --set full` report into the per-launch summary kept under profiles/:
   python scripts/ncu_summary.py gpurun_out/prof_tiles.ncu-rep tiles > profiles/rNN_ncu_tiles_summary.csv
Row 0 = metric names, row 1 = units, then one row per captured launch (ncu -i ... --page raw --csv)."""
import csv
import io
import subprocess
import sys

COLS = {
    "tiles": ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
              "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"],
    "pointwise": ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
                  "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                  "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                  "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
                  "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
                  "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
                  "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"],
}


def main():
    rep, kind = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(c) for c in COLS[kind] if c in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()
