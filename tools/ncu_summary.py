"""Condense an .ncu-rep into the text summary kept under profiles/ (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "header line" > profiles/rNN_ncu_<kernel>.txt
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_srcunit_tex_aperture_sysmem_op_write.sum", "pcie__write_bytes.sum", "pcie__read_bytes.sum",
]


def main():
    rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(names)}
    if header:
        print(header)
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']]}")
        for k in KEEP:
            if k in col and r[col[k]] != "":
                print(f"{k:<85} {r[col[k]]} {units[col[k]]}")


if __name__ == "__main__":
    main()
