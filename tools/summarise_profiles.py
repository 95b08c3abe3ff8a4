"""Turns the raw artefacts of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
G = os.path.join(ROOT, 'gpurun_out')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r1'


def last_json(path):
    lines = [l for l in open(path) if l.startswith('{')]
    return json.loads(lines[-1]) if lines else None


def launches():
    rows = list(csv.reader(open(os.path.join(G, 'launches.csv'))))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    d = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > mv:
            try:
                d[r[kn]].append(float(r[mv].replace(',', '')))
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    out = ['ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 (bf16 mode)',
           'per-launch times are cold-cache and serialised: compare SHARES with the live CUDA-event numbers of bench.py', '',
           '%-78s %5s %10s %7s' % ('kernel', 'n', 'mean us', 'share')]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append('%-78s %5d %10.1f %6.1f%%' % (k[:78], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    return '\n'.join(out) + '\n'


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu.sum',
        'sm__cycles_elapsed.avg']


def ncu_raw(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ['ncu --set full --clock-control none --import-source on  (%s)' % os.path.basename(rep), '']
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        if name in seen:
            continue
        seen.add(name)
        out.append('== ' + name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append('  %-66s %16s %s' % (k, r[i], units[i]))
        out.append('')
    return '\n'.join(out)


def main():
    os.makedirs(OUT, exist_ok=True)
    for src, dst in (('bench_bf16.log', 'bench_bf16.json'), ('bench_reference.log', 'bench_reference.json')):
        p = os.path.join(G, src)
        if os.path.exists(p):
            j = last_json(p)
            if j:
                json.dump(j, open(os.path.join(OUT, '%s_%s' % (TAG, dst)), 'w'), indent=1)
    if os.path.exists(os.path.join(G, 'launches.csv')):
        open(os.path.join(OUT, '%s_launches_summary.txt' % TAG), 'w').write(launches())
        subprocess.run(['cp', os.path.join(G, 'launches.csv'), os.path.join(OUT, '%s_launches.csv' % TAG)])
    rep = os.path.join(G, 'prof_edge_ws.ncu-rep')
    if os.path.exists(rep):
        open(os.path.join(OUT, '%s_ncu_full_summary.txt' % TAG), 'w').write(ncu_raw(rep))
    for r in (1, 2, 3):
        p = os.path.join(G, 'trace_role%d.log' % r)
        if os.path.exists(p):
            subprocess.run(['cp', p, os.path.join(OUT, '%s_ws_trace_role%d.txt' % (TAG, r))])


if __name__ == '__main__':
    main()
