"""mbarrier / named-barrier wait sites of one kernel with their stall samples
(input: ncu -i rep --page source --csv --print-source sass [--launch-skip i --launch-count 1])."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot)
for i, r in enumerate(data):
    s = r[ix['Source']]
    if 'TRYWAIT' in s or 'BAR.SYNC' in s or 'BAR.' in s:
        smp = sum(int(data[j][ix['# Samples']] or 0) for j in range(i, min(i + 4, len(data))))
        print('%5d %6d smp (%4.1f%%) %8s exe  %s' % (i, smp, smp * 100.0 / tot, r[ix['Instructions Executed']], s[:100]))
