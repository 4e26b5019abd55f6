"""Turn an .ncu-rep (ncu --set full) into the small per-kernel summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_r1_step.ncu-rep profiles/r1_step_f32_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = [f"# ncu --set full summary of `{rep}`", "",
             "Per-launch values (cold-cache, serialised replays; compare shares, not absolutes).", ""]
    for n, r in enumerate(data):
        lines += [f"## launch {n}: `{r[ki]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} | {r[i]} | {units[i]} |")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
