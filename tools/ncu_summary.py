"""Print the handful of ncu metrics we track from a .ncu-rep (raw page)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.sum.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
 'smsp__average_warp_latency_issue_stalled_short_scoreboard','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:95s} {units[i]:12s} {r[i]}")
    print("-" * 40)
