"""Counts of SASS mnemonics per kernel of the built library (evidence that the hot kernels are tcgen05 / TMA code):
    python tools/sass_counts.py [path/to/libb200ns.so] > profiles/rNN_sass_instruction_counts.txt
UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / .st,
SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0), MUFU = special-function unit, D-fp64 = DADD / DMUL / DFMA."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'diffusion-tts_b200', 'libb200ns.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
cols = ['UTCHMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'SYNCS', 'HMMA', 'MUFU']
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(1)
        c = counts[cur]
        c['instr'] += 1
        base = op.split('.')[0]
        if base in cols:
            c[base] += 1
        if base in ('DADD', 'DMUL', 'DFMA'):
            c['D-fp64'] += 1
names = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
print(f'cuobjdump -sass {os.path.relpath(so, ROOT)}  (sm_100a; counts of SASS mnemonics per kernel; tools/sass_counts.py)\n')
print(f'{"kernel":88s}' + ''.join(f'{h:>8s}' for h in ['instr'] + cols + ['D-fp64']))
tot = collections.Counter()
for (mangled, c), name in sorted(zip(counts.items(), names), key=lambda t: -t[0][1]['instr']):
    name = re.sub(r'\(.*', '', name)
    print(f'{name[:87]:88s}' + ''.join(f'{c[h]:8d}' for h in ['instr'] + cols + ['D-fp64']))
    tot.update(c)
print(f'{"TOTAL (" + str(len(counts)) + " kernels)":88s}' + ''.join(f'{tot[h]:8d}' for h in ['instr'] + cols + ['D-fp64']))
