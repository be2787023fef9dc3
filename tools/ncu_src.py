"""Developer tool: summarise an `ncu --page source --csv` dump (one kernel section): opcode mix and the
instructions with most stall samples.   python tools/ncu_src.py file.csv <first_line> <last_line>"""
import collections
import csv
import sys

path, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lines = open(path).read().splitlines()[lo - 1:hi]
rows = list(csv.reader(lines))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
iS, iE, iP, iA = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address')
tot = sum(int(r[iE]) for r in data)
ts = sum(int(r[iP]) for r in data)
print("instructions", tot, "samples", ts)
byop, samp = collections.Counter(), collections.Counter()
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    byop[op] += int(r[iE]); samp[op] += int(r[iP])
for op, n in byop.most_common(18):
    print("%-10s %10d %5.1f%%  samples %5.1f%%" % (op, n, 100 * n / tot, 100 * samp[op] / max(ts, 1)))
print("most stalled instructions:")
for r in sorted(data, key=lambda r: -int(r[iP]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 24]:
    print(r[iA][-5:], "%9s %5s  %s" % (r[iE], r[iP], r[iS][:110]))
