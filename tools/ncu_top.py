"""Top stall-sampled SASS instructions of one kernel from `ncu --page source --csv` output."""
import csv, sys
allrows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name']
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = allrows[starts[sec]:(starts[sec + 1] if sec + 1 < len(starts) else len(allrows))]
print('sections', len(starts), '->', rows[0][1][:60])
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[i_s]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for i, h in stall:
        agg[h] = agg.get(h, 0) + int(r[i])
print(sorted(agg.items(), key=lambda t: -t[1])[:8])
for n, r in sorted(enumerate(data), key=lambda t: -int(t[1][i_s]))[:n_top]:
    st = sorted(((int(r[i]), h) for i, h in stall), reverse=True)[:2]
    print(n, r[i_s], r[i_ex], r[i_src].strip()[:80], st)
