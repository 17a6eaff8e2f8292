import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; units = rows[1]; data = rows[2:]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','SM_A.TriageCompute.l1tex__throughput.avg.pct_of_peak_sustained_elapsed','LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active',
 'smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_mio_throttle','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_not_selected','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_lg_throttle','smsp__pcsamp_warps_issue_stalled_dispatch_stall','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_branch_resolving','smsp__pcsamp_warps_issue_stalled_membar','smsp__pcsamp_warps_issue_stalled_imc_miss','smsp__pcsamp_warps_issue_stalled_sleeping','smsp__pcsamp_warps_issue_stalled_tex_throttle','smsp__pcsamp_warps_issue_stalled_drain',
 'lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','sm__cycles_elapsed.avg','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed']
idx = {h:i for i,h in enumerate(hdr)}
for r in data:
    print('='*100)
    for w in want:
        if w in idx:
            i = idx[w]
            print(f"  {w:75s} {r[i]:>22s} {units[i]}")
if len(sys.argv) > 2:
    print([h for h in hdr if sys.argv[2] in h])
