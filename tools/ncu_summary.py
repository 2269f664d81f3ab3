"""Summarises an .ncu-rep (read here, no GPU needed) into the handful of numbers the roofline notes cite."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "local_load", "local_store",
    "smsp__inst_executed_op_local", "l1tex__t_bytes_pipe_lsu_mem_local",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print("==", r[name_i][:110])
        for i, h in enumerate(hdr):
            if any(h == k or (k in h and len(k) < 24) for k in KEYS):
                print(f"  {h:75s} {r[i]:>16s} {units[i]}")
        stalls = [(float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
        print("  stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
