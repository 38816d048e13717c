"""Summarise the mbarrier wait sites (and a few other blocking instructions) of one kernel in an ncu SASS CSV.
usage: python scripts/ncu_waits.py sass.csv [kernel index]"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
heads = [i for i, x in enumerate(r) if x and x[0] == 'Address']
hdr = r[heads[which]]
rows = []
for x in r[heads[which] + 1:]:
    if not x or not x[0].startswith('0x'):
        break
    rows.append(x)
isamp = hdr.index('# Samples'); ia = hdr.index('Source'); iex = hdr.index('Instructions Executed')
total = sum(int(x[isamp]) for x in rows)
print('total samples', total, 'instructions', len(rows))
tw = 0
for i, x in enumerate(rows):
    if 'TRYWAIT' in x[ia]:
        s = int(rows[i + 1][isamp]) + int(x[isamp])
        tw += s
        print('%5d %6d (%4.1f%%) exec %9s  %s' % (i, s, 100.0 * s / total, x[iex], x[ia].strip()[:70]))
print('mbarrier waits: %d samples (%.1f%%)' % (tw, 100.0 * tw / total))
for key in ['BAR.SYNC', 'ATOMG', 'MEMBAR', 'LDG', 'NANOSLEEP', 'UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'LDGSTS']:
    s = e = 0
    for i, x in enumerate(rows):
        if key in x[ia]:
            s += int(x[isamp]) + (int(rows[i + 1][isamp]) if i + 1 < len(rows) else 0)
            e += int(x[iex])
    print('%-10s samples(+next) %6d  executed %d' % (key, s, e))
