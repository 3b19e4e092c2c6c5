"""Print the key metrics of every kernel in an ncu report: python scripts/ncu_keys.py report.ncu-rep [name-filter]"""
import csv, subprocess, sys
rep = sys.argv[1]; flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); h = r[0]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct']
for row in r[2:]:
    name = row[h.index('Kernel Name')]
    if flt and flt not in name: continue
    print('==', name[:60])
    for k, v, u in zip(h, row, r[1]):
        if k in keys: print('   %-70s %s %s' % (k, v, u))
    st = [(float(v), k) for k, v in zip(h, row) if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k]
    print('   stalls:', ', '.join('%s %.2f' % (k.split('issue_stalled_')[1].split('_per_')[0], v) for v, k in sorted(st, reverse=True)[:6]))
