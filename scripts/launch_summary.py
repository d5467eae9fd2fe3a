"""Summarise an ncu launch list (csv of gpu__time_duration.sum) into a per-kernel share table.
usage: python scripts/launch_summary.py gpurun_out/x.csv [steps]"""
import csv, sys, re, collections
path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    us = v / 1000.0 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1000.0)
    name = r['Kernel Name']
    name = re.sub(r'\(.*$', '', name)[:110]
    tot[name][0] += us
    tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print("total serialised us: %.1f  (per step %.1f)" % (total, total / steps))
print("| share | us/launch | launches/step | kernel |\n|---|---|---|---|")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
    print("| %.2f%% | %.1f | %.1f | `%s` |" % (100 * us / total, us / n, n / steps, name))
own = sum(v[0] for k, v in tot.items() if 'feta::' in k)
lib = sum(v[0] for k, v in tot.items() if 'cutlass' in k or 'cublas' in k or 'gemm' in k.lower())
print("own %.1f%%, library GEMM %.1f%%" % (100 * own / total, 100 * lib / total))
