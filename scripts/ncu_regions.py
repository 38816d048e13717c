"""Aggregate ncu source-page (SASS) stall samples into contiguous code regions.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass --launch-skip N --launch-count 1 > sass.csv
       python scripts/ncu_regions.py sass.csv [chunk]"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 150
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
heads = [i for i, x in enumerate(r) if x and x[0] == 'Address']
print('kernels in file:', [r[i - 1][1][:40] for i in heads])
hi = heads[which]
hdr = r[hi]
rows = []
for x in r[hi + 1:]:
    if not x or not x[0].startswith('0x'):
        break
    rows.append(x)
ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(x[isamp]) for x in rows); totex = sum(int(x[iex]) for x in rows)
print('total samples', tot, 'instrs', len(rows), 'executed', totex)
for b in range(0, len(rows), CH):
    ch = rows[b:b + CH]
    s = sum(int(x[isamp]) for x in ch); ex = sum(int(x[iex]) for x in ch)
    ops = {}
    for x in ch:
        t = x[ia].split()
        op = t[1] if t[0].startswith('@') else t[0]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda t: -t[1])[:5]
    st = {}
    for x in ch:
        for i in stall_cols:
            v = int(x[i]) if x[i] else 0
            if v: st[hdr[i][6:]] = st.get(hdr[i][6:], 0) + v
    tops = sorted(st.items(), key=lambda t: -t[1])[:4]
    print('%5d samp %5.1f%% exec %5.1f%%' % (b, 100 * s / tot, 100 * ex / totex), top, tops)
