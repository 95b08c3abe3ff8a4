"""Per-source-line executed warp instructions and stall samples
(input: ncu -i rep --page source --csv --print-source cuda,sass [--launch-skip i --launch-count 1])."""
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
data, cur, hdr = [], '', None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = os.path.basename(r[1])
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] != '':
        data.append((cur, r))
i_ins, i_smp = hdr.index('Instructions Executed'), hdr.index('# Samples')
tot_i = sum(int(r[i_ins] or 0) for _, r in data)
tot_s = sum(int(r[i_smp] or 0) for _, r in data)
print('warp instructions', tot_i, 'samples', tot_s)
for f, r in data:
    ins, smp = int(r[i_ins] or 0), int(r[i_smp] or 0)
    if ins * 100.0 / tot_i >= thr or smp * 100.0 / tot_s >= thr:
        print('%-16s %4s %5.1f%% ins %5.1f%% smp | %s' % (f[:16], r[0], ins * 100.0 / tot_i, smp * 100.0 / tot_s, r[1].strip()[:110]))
