"""Condense an `ncu --set full` report (raw page CSV on stdin) to the columns the design doc cites."""
import csv, sys
cols = ['ID', 'Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum']
r = list(csv.reader(sys.stdin))
h = r[0]
idx = [h.index(c) for c in cols if c in h]
w = csv.writer(sys.stdout)
for row in r:
    w.writerow([row[i] for i in idx])
