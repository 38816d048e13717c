"""Top sampled SASS instructions of one kernel in an ncu source-page CSV, with a few neighbours for context.
usage: python scripts/ncu_top.py sass.csv <kernel index> [N]"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]); N = int(sys.argv[3]) if len(sys.argv) > 3 else 12
heads = [i for i, x in enumerate(r) if x and x[0] == 'Address']
hdr = r[heads[which]]
rows = []
for x in r[heads[which] + 1:]:
    if not x or not x[0].startswith('0x'):
        break
    rows.append(x)
isamp = hdr.index('# Samples'); ia = hdr.index('Source'); iex = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(x[isamp]) for x in rows)
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][isamp]))[:N]
for i in order:
    x = rows[i]
    st = {hdr[j][6:]: int(x[j]) for j in stall_cols if x[j] and int(x[j])}
    print('== #%d  %.1f%% of samples, executed %s: %s' % (i, 100 * int(x[isamp]) / tot, x[iex], sorted(st.items(), key=lambda t: -t[1])[:2]))
    for k in range(max(0, i - 6), min(len(rows), i + 2)):
        print('     %4d %6s  %s' % (k, rows[k][isamp], rows[k][ia].strip()[:90]))
