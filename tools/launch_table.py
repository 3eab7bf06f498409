"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.OrderedDict(); tot = 0
for row in csv.DictReader(lines):
    name = row['Kernel Name']
    t = float(row['Metric Value'].replace(',', ''))
    if row['Metric Unit'] == 'ns': t /= 1000
    elif row['Metric Unit'] == 'ms': t *= 1000
    m = re.search(r'gemm_bf16_kernel<([^>]*)>', name)
    key = ('gemm<' + m.group(1).replace('(int)', '').replace('(bool)', '') + '>') if m else re.sub(r'\(.*', '', name)[-44:]
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t/1000:8.3f} ms {100*t/tot:5.1f}%  n={n:4d} avg {t/n:8.1f} us  {k}")
print(f"total {tot/1000:.3f} ms over {sum(v[0] for v in agg.values())} launches")
