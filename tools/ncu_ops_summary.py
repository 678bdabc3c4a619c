import csv, sys
cols = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size']
r = list(csv.reader(sys.stdin)); h = r[0]
idx = [h.index(c) for c in cols if c in h]
w = csv.writer(sys.stdout)
seen = set()
for i, row in enumerate(r):
    if i >= 2:
        key = row[h.index('Kernel Name')]
        if key in seen: continue
        seen.add(key)
    w.writerow([row[j][:90] for j in idx])
