"""Join an ncu launch list (gemm kernels, in launch order) with the B200_GEMM_TRACE lines of the same run."""
import collections, csv, re, sys
trace = [l.split()[1:] for l in open(sys.argv[2]) if l.startswith('GEMMTRACE')]
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
durs = []
for row in csv.DictReader(lines):
    if 'gemm_tcgen05' in row['Kernel Name']:
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        durs.append(v / 1e3 if u.startswith('n') else (v * 1e3 if u.startswith('m') else v))
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0       # gemm launches skipped by ncu -s (counted in gemm launches)
trace = trace[skip:skip + len(durs)]
agg = collections.defaultdict(lambda: [0, 0.0])
for t, d in zip(trace, durs):
    agg[' '.join(t)][0] += 1
    agg[' '.join(t)][1] += d
tot = sum(v[1] for v in agg.values())
print(f'{len(durs)} gemm launches, {tot:.0f} us')
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, K = map(int, k.split()[:3])
    print(f'{us:9.0f} us {100*us/tot:5.1f}%  n={n:3d} avg {us/n:7.1f} us  {2.0*M*N*K*n/us*1e-6:7.1f} TF/s  {k}')
