"""Developer tool (no GPU needed): per-kernel SASS opcode counts of libb200e2tts.so — evidence that the hot kernels are tcgen05 / TMEM / TMA
code (B200_PROFILING.md table: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor load/store -> UTMALDG / UTMASTG,
cp.async.bulk -> UBLKCP, mma.sync -> HMMA, packed fp32x2 -> FFMA2).  usage: python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'e2-tts-pytorch_b200', 'libb200e2tts.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
OPS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'HMMA', 'REDG', 'FFMA2']
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(.*', '', cur).replace('void ', '')
        counts[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(1)
        base = op.split('.')[0]
        counts[cur][base] += 1
        if op.startswith('UTCHMMA.2CTA') or '.2CTA' in op and base == 'UTCHMMA':
            counts[cur]['UTCHMMA.2CTA'] += 1
print('SASS evidence for libb200e2tts.so (final round-2 build): opcode counts per kernel from `cuobjdump -sass`')
print('tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, cp.async.bulk.tensor load / store -> UTMALDG / UTMASTG,')
print('cp.async.bulk -> UBLKCP, mma.sync (legacy cross-check kernels only) -> HMMA, red.global.add.v4.f32 -> REDG, fp32x2 FMA -> FFMA2\n')
print(f'{"kernel":84s}' + ''.join(f'{o:>13s}' for o in OPS))
for k, c in counts.items():
    if any(c[o] for o in OPS):
        print(f'{k[:83]:84s}' + ''.join(f'{c[o]:13d}' for o in OPS))
