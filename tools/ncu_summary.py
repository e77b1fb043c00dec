"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
]


SUBSTR = ["sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
          "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for row in rows[2:]:
        print("=" * 100)
        print(row[name_col][:90], "| grid", row[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for h, u, v in zip(hdr, units, row):
            if h in KEYS or any(k in h for k in SUBSTR):
                print(f"  {h:90s} {v:>16s} {u}")


if __name__ == "__main__":
    main(sys.argv[1])
