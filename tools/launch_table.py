"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list per kernel: time share, launches, average duration and (when captured) DRAM traffic."""
import collections, csv, json, os, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith('--') else 25
json_out = sys.argv[sys.argv.index('--json') + 1] if '--json' in sys.argv else None  # totals for bench.py's roofline.traffic
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
SCALE = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6,
         'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
agg = collections.OrderedDict()
tot_t = tot_b = 0.0
ids = set()
for row in csv.DictReader(lines):
    name = row['Kernel Name']
    val = float(row['Metric Value'].replace(',', '')) * SCALE.get(row['Metric Unit'], 1.0)
    m = re.search(r'gemm_bf16_kernel<([^>]*)>', name)
    key = ('gemm<' + m.group(1).replace('(int)', '').replace('(bool)', '') + '>') if m else re.sub(r'\(.*', '', name)[-44:]
    a = agg.setdefault(key, [set(), 0.0, 0.0])
    a[0].add(row['ID'])
    ids.add(row['ID'])
    if row['Metric Name'] == 'gpu__time_duration.sum':
        a[1] += val; tot_t += val
    elif row['Metric Name'].startswith('dram__bytes'):
        a[2] += val; tot_b += val
for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    extra = f"  dram {b/1e6:9.1f} MB ({b/1e3/max(t,1e-9):6.0f} GB/s)" if tot_b else ""
    print(f"{t/1000:8.3f} ms {100*t/tot_t:5.1f}%  n={len(n):4d} avg {t/len(n):8.1f} us  {k}{extra}")
print(f"total {tot_t/1000:.3f} ms over {len(ids)} launches" + (f"; DRAM traffic {tot_b/1e9:.2f} GB = {tot_b/1e3/tot_t:.0f} GB/s "
      f"averaged over the summed kernel time" if tot_b else ""))
if json_out:
    json.dump({"source": os.path.basename(path), "launches": len(ids), "dram_bytes": tot_b, "kernel_time_us": tot_t,
               "note": "sum over all launches of one DSFVT train step (ncu, cold-cache, serialised): "
                       "dram__bytes_read.sum + dram__bytes_write.sum"}, open(json_out, "w"), indent=1)
