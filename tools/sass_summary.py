"""Counts the Blackwell-native / legacy tensor-path SASS mnemonics per kernel of the in-tree library.

    python tools/sass_summary.py [out.txt]      (default: profiles/r2_sass_summary.txt)

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP / UTMALDG = TMA bulk / tensor copies, UTCBAR = tcgen05.commit,
HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async  (B200_PROFILING.md "What proves a Blackwell-native kernel")."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'shapemol_b200', 'libshapemol_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'HMMA', 'LDGSTS', 'FFMA2', 'MUFU']


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r2_sass_summary.txt')
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            counts[cur]['total'] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    demangle = subprocess.run(['c++filt'] + list(counts), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (name, c), dm in zip(counts.items(), demangle):
        short = re.sub(r'\(anonymous namespace\)::|smb::', '', dm)
        short = re.sub(r'\(.*\)$', '', short)[:60]
        rows.append((short, c))
    with open(out_path, 'w') as f:
        f.write('cuobjdump -sass shapemol_b200/libshapemol_b200.so: SASS mnemonic counts per kernel (sm_100a)\n')
        f.write('%-60s %7s ' % ('kernel', 'instrs') + ' '.join('%7s' % k for k in KEYS) + '\n')
        for short, c in sorted(rows):
            f.write('%-60s %7d ' % (short, c['total']) + ' '.join('%7d' % c[k] for k in KEYS) + '\n')
        tot = collections.Counter()
        for _, c in rows:
            tot.update(c)
        f.write('%-60s %7d ' % ('TOTAL', tot['total']) + ' '.join('%7d' % tot[k] for k in KEYS) + '\n')
    print(open(out_path).read())


if __name__ == '__main__':
    main()
