"""profiles/r2_conv_traffic.json from an ncu metrics CSV (one eager generator forward, conv_tc launches only):

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none --profile-from-start off -k regex:conv_tc --csv --log-file gpurun_out/conv_metrics.csv python tools/fwd_once.py
    python tools/ncu_traffic.py gpurun_out/conv_metrics.csv <algorithmic bytes per forward> > profiles/r2_conv_traffic.json

Algorithmic bytes of one conv launch = fp32 inputs read once (all K segments) + packed weights read once + fp32 output written once
(+ residual read once); the engine reports the sum over its conv launches (GeneratorEngine.conv_bytes)."""
import csv, json, sys, collections
path = sys.argv[1]
alg_bytes = float(sys.argv[2]) if len(sys.argv) > 2 else None
lines = [l for l in open(path) if l.startswith('"')]
per = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row['ID']
    d = per.setdefault(k, {'name': row['Kernel Name']})
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    m = row['Metric Name']
    if m.startswith('dram__bytes'):
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        d[m] = v * mult
    elif m == 'gpu__time_duration.sum':
        d[m] = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
    else:
        d[m] = v
n = len(per)
rd = sum(d.get('dram__bytes_read.sum', 0) for d in per.values())
wr = sum(d.get('dram__bytes_write.sum', 0) for d in per.values())
us = sum(d.get('gpu__time_duration.sum', 0) for d in per.values())
tp = [d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') for d in per.values()]
tp = [t for t in tp if t is not None]
wtp = sum(d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0) * d.get('gpu__time_duration.sum', 0) for d in per.values()) / max(us, 1e-9)
out = {'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (+ duration, tensor pipe) over every conv_tc launch of one eager '
                 'generator forward (tools/fwd_once.py, batch 64, CIFAR-10 NCSN++), cold-cache serialised launches; profiles/r2_conv_metrics.csv',
       'launches': n, 'dram_bytes_read_per_forward': rd, 'dram_bytes_write_per_forward': wr,
       'dram_bytes_per_launch': (rd + wr) / max(n, 1), 'ncu_us_per_forward': us,
       'tensor_pipe_active_pct_time_weighted': wtp,
       'algorithmic_bytes_per_launch': (alg_bytes / n) if alg_bytes else None,
       'algorithmic_bytes_per_forward': alg_bytes}
print(json.dumps(out, indent=1))
