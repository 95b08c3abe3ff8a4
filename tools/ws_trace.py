"""Per-tile role timeline of CTA 0 of the LAST warp-specialised edge kernel of one forward (SMB_WS_DBG must include 16).
Prints, per tile, the clock64 stamps relative to the first event, and the mean per-tile period of every event."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
from conftest import load_golden  # noqa: E402
from test_gpu_parity import build_model, batch_of  # noqa: E402
from shapemol_b200 import _lib  # noqa: E402

NAMES = ['P arrive', 'G1 issue', 'LN start', 'LN arrive', 'G2 issue', 'E2 start', 'E2 end', 'G1 a1full', 'G1 mma\'d', 'G1 commit', 'G1 a1free', 'G1 loaded', 'G1 dfree', 'P prewait', 'P postwait', 'CV qb', 'CV mfree', 'CV bm', 'CV folded', 'CV staged']


def main():
    fx = load_golden('forward_k32_eval.pt')
    m = build_model(fx, 'bf16', training=False)
    g = torch.Generator().manual_seed(11)
    B = 3000
    sizes = torch.randint(9, 28, (B,), generator=g).tolist()
    if os.environ.get('WS_TRACE_N'):
        sizes = [int(os.environ['WS_TRACE_N'])] * B
    N = sum(sizes)
    pos = torch.randn(N, 3, generator=g).cuda() * 2
    v = torch.randint(0, 15, (N,), generator=g).cuda()
    shape = (0.07 * torch.randn(B, 32, 3, generator=g)).cuda()
    t = torch.randint(0, 1000, (B,), generator=g).cuda()
    batch = batch_of(sizes)
    for _ in range(2):
        m(pos, v, batch, shape, time_step=t)
    torch.cuda.synchronize()
    lib = _lib.load()
    buf = np.zeros((20, 128), dtype=np.int64)
    lib.smb_debug_ws_trace.argtypes = [C.c_void_p]
    rc = lib.smb_debug_ws_trace(buf.ctypes.data)
    assert rc == 0, rc
    nt = int((buf[0] > 0).sum())
    t0 = buf[buf > 0].min()
    rel = np.where(buf > 0, buf - t0, -1)
    order = [13, 14, 0, 7, 12, 1, 8, 9, 10, 11, 2, 3, 15, 19, 18, 16, 17, 4, 5, 6]
    print('tiles traced', nt)
    print('tile ' + ' '.join('%10s' % NAMES[e] for e in order))
    for i in list(range(0, 6)) + list(range(40, 48)):
        if i < nt:
            print('%4d ' % i + ' '.join('%10d' % rel[e][i] for e in order))
    a, b = 20, min(nt, 80) - 1
    print('mean period (tiles %d..%d):' % (a, b), ' '.join('%s %.0f' % (NAMES[e], (buf[e][b] - buf[e][a]) / (b - a)) for e in order))
    for (x, y) in ((1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (15, 19), (19, 18), (18, 16), (17, 15), (16, 17), (17, 4), (13, 14), (14, 0), (0, 13), (0, 1), (7, 12), (12, 1), (1, 8), (8, 9), (9, 10), (10, 11)):
        if buf[x][a:b].min() <= 0 or buf[y][a:b].min() <= 0:
            continue
        d = (buf[y][a:b] - buf[x][a:b])
        print('%-10s -> %-10s mean %6.0f  min %6d  max %6d' % (NAMES[x], NAMES[y], d.mean(), d.min(), d.max()))


if __name__ == '__main__':
    main()
