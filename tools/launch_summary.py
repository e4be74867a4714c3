"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (time share, count, mean)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit.startswith('n') else (v * 1e3 if unit.startswith('m') else v)
    name = re.sub(r'<.*', '', re.sub(r'\(.*', '', row['Kernel Name'])).replace('void ', '')
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f'total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{t:10.0f} us {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:8.1f} us  {k}')
