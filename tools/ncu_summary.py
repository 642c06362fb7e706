#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` into one block per launch with the metrics the design notes quote."""
import csv
import sys

KEYS = [
    ("time_us", "gpu__time_duration.sum"), ("grid", "Grid Size"), ("block", "Block Size"),
    ("regs", "launch__registers_per_thread"), ("smem_dyn_B", "launch__shared_mem_per_block_dynamic"),
    ("waves", "launch__waves_per_multiprocessor"), ("occ_theo_%", "sm__maximum_warps_per_active_cycle_pct"),
    ("occ_achv_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
    ("dram_%peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_%peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_%peak", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("lts_%peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warp_insts", "smsp__inst_executed.sum"), ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("ipc_active", "sm__inst_executed.avg.per_cycle_active"), ("issue_active_%", "smsp__issue_active.avg.pct"),
    ("ld_sectors/req", "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio"),
    ("st_sectors/req", "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio"),
    ("smem_bank_conf_ld", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("smem_bank_conf_st", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
    ("local_ld", "smsp__inst_executed_op_local_ld.sum"), ("local_st", "smsp__inst_executed_op_local_st.sum"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("stall_not_sel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("pipe_alu_%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe_fma_%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe_lsu_%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
]


def main(path, only=None):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[h], rows[h + 1]
    for r in rows[h + 2:]:
        if len(r) != len(hdr):
            continue
        name = r[hdr.index("Kernel Name")].split("(")[0]
        if only and only not in name:
            continue
        print("== %s" % name)
        parts = []
        for label, key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                parts.append("%s=%s%s" % (label, v, (" " + units[i]) if units[i] and label in ("dram_rd", "dram_wr") else ""))
        for i in range(0, len(parts), 6):
            print("   " + "  ".join(parts[i:i + 6]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
