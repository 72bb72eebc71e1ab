"""Top warp-stall sample locations (SASS) of an ncu report's source page.
    ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N]
"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
si = hdr.index('# Samples')
data = []
for k, r in enumerate(rows[hdr_i + 1:]):
    try:
        data.append((float(r[si] or 0), k, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print('total samples', tot, ' instructions', len(data))
for v, k, r in sorted(data, key=lambda t: -t[0])[:n]:
    print('%7.0f %5.1f%%  #%-5d %s   [exec %s]' % (v, 100 * v / tot, k, r[1].strip()[:100], r[5]))
