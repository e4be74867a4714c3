"""DRAM traffic of the GEMM kernel family per step from an ncu launch list:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 -s <skip> -c <n> \
        --csv --log-file gpurun_out/gemm_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph
    python tools/gemm_traffic.py gpurun_out/gemm_traffic.csv <launches_per_step> cfg2   ->  profiles/r2_gemm_traffic.json

bench.py reads that file for `roofline.traffic` (bytes moved through DRAM by all GEMM launches of one step)."""
import collections, csv, json, os, sys
path, per_step, key = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = collections.defaultdict(dict)
for row in csv.DictReader(l for l in open(path) if not l.startswith('==')):
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit'].lower()
    if 'byte' in u:
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    rows[int(row['ID'])][row['Metric Name']] = v
ids = sorted(rows)
ids = ids[len(ids) % per_step:] if per_step and len(ids) >= per_step else ids      # whole steps only (drop a leading partial step)
nsteps = max(1, len(ids) // per_step) if per_step else 1
rd = sum(rows[i].get('dram__bytes_read.sum', 0.0) for i in ids) / nsteps
wr = sum(rows[i].get('dram__bytes_write.sum', 0.0) for i in ids) / nsteps
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r2_gemm_traffic.json')
data = json.load(open(out)) if os.path.exists(out) else {}
data[key] = rd + wr
data[key + '_detail'] = dict(read_bytes=rd, write_bytes=wr, gemm_launches=len(ids), steps=nsteps, source=os.path.basename(path))
json.dump(data, open(out, 'w'), indent=1)
print(key, 'GEMM DRAM traffic per step: %.2f GB read + %.2f GB write over %d launches / %d step(s)' % (rd / 1e9, wr / 1e9, len(ids), nsteps))
