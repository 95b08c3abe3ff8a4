"""Top stall sites of one kernel from `ncu -i rep --page source --csv --print-source sass` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
texec = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
print('total samples', tot, 'warp instructions', texec)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print('%5d %6s smp %8s exe  %-70s %s' % (i, r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']][:70],
                                             ' '.join('%s=%d' % (c, n) for n, c in st if n)))
