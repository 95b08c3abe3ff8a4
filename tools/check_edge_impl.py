"""Prints the bf16-mode forward errors against the reference fixtures (and the oracle on a large ragged
batch) for the edge implementation selected by the environment (SMB_EDGE_LEGACY=1 -> mma.sync kernels)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
from conftest import load_golden  # noqa: E402
from test_gpu_parity import build_model, batch_of  # noqa: E402


def errs(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


def main():
    tag = 'legacy' if os.environ.get('SMB_EDGE_LEGACY') else 'ws'
    if os.environ.get('SMB_NODE_LEGACY'):
        tag += '_nl'
    for name in ('k32_train', 'k32_eval', 'k8_train', 'k8_eval', 'tiny_train'):
        fx = load_golden('forward_%s.pt' % name)
        for prec in ('bf16',):
            m = build_model(fx, prec)
            out = m(fx['pos'].cuda(), fx['v'].cuda(), batch_of(fx['sizes']), fx['shape'].cuda(), time_step=fx['t'].cuda())
            torch.cuda.synchronize()
            e = {k: errs(out[k], fx[r]) for k, r in (('pred_ligand_pos', 'pred_pos'), ('pred_ligand_h', 'pred_h'), ('pred_ligand_v', 'pred_v'))}
            print(tag, name, prec, ' '.join('%s max %.2e l2 %.2e' % (k[12:], v[0], v[1]) for k, v in e.items()), flush=True)
    # large ragged batch: finite, and timing
    fx = load_golden('forward_k32_eval.pt')
    m = build_model(fx, 'bf16', training=False)
    g = torch.Generator().manual_seed(11)
    B = 3000
    sizes = torch.randint(1, 28, (B,), generator=g).tolist()
    N = sum(sizes)
    pos = torch.randn(N, 3, generator=g).cuda() * 2
    v = torch.randint(0, 15, (N,), generator=g).cuda()
    shape = (0.07 * torch.randn(B, 32, 3, generator=g)).cuda()
    t = torch.randint(0, 1000, (B,), generator=g).cuda()
    batch = batch_of(sizes)
    out = m(pos, v, batch, shape, time_step=t)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        out = m(pos, v, batch, shape, time_step=t)
    torch.cuda.synchronize()
    print(tag, 'ragged 1..27: finite', all(bool(torch.isfinite(x).all()) for x in out.values()), 'ms/forward %.2f' % ((time.time() - t0) / 3 * 1e3))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    torch.save({k: x.cpu() for k, x in out.items()}, os.path.join(ROOT, 'gpurun_out', 'ragged_%s.pt' % tag))
    for other in ('legacy', 'ws', 'ws_nl'):
        f = os.path.join(ROOT, 'gpurun_out', 'ragged_%s.pt' % other)
        if other != tag and os.path.exists(f):
            ref = torch.load(f)
            print(tag, 'vs', other, ' '.join('%s max %.2e l2 %.2e' % ((k[12:],) + errs(out[k], ref[k])) for k in out), flush=True)


if __name__ == '__main__':
    main()
