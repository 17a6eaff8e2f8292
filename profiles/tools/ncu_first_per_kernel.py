"""One block of key metrics per distinct (kernel name, grid size) of an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
seen = set()
for r in data:
    key = (r[idx['Kernel Name']], r[idx['launch__grid_size']])
    if key in seen:
        continue
    seen.add(key)
    print('=' * 110)
    print(key[0][:150], ' grid', key[1])
    for w in want:
        if w in idx:
            print(f"  {w:80s} {r[idx[w]]:>18s} {units[idx[w]]}")
