"""Developer tool: top SASS instructions by warp-stall samples from an .ncu-rep (source page), per kernel."""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
lines = out.splitlines()
blocks, cur = [], None
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        cur = [ln]; blocks.append(cur)
    elif cur is not None:
        cur.append(ln)
for b in blocks:
    name = next(csv.reader([b[0]]))[1]
    rows = list(csv.reader(io.StringIO('\n'.join(b[1:]))))
    hdr = rows[0]
    i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    data = [(int(r[i_s] or 0), idx, r[i_src].strip(), int(r[i_ex] or 0)) for idx, r in enumerate(rows[1:]) if len(r) > i_ex]
    tot = sum(d[0] for d in data) or 1
    print('==', name[:100], 'samples', tot)
    for smp, idx, src, ex in sorted(data, reverse=True)[:top]:
        print(f'  {100.0 * smp / tot:5.1f}%  #{idx:5d}  exec={ex:9d}  {src[:110]}')
