"""Step time of the generic-shape fp32 path (hidden 256, k = 48, 60-atom molecules: the BASELINE configs[4] shape).
    python tools/bench_generic.py [--mols 256]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mols', type=int, default=256)
    ap.add_argument('--steps', type=int, default=3)
    args = ap.parse_args()
    from conftest import load_golden, golden_weights
    from test_host_cpu import make_dropin
    from shapemol_b200.engine import Sampler
    fx = load_golden('forward_k48_h256_train.pt')
    m, _ = make_dropin(knn=48, hidden_dim=256, n_heads=16)
    m.load_state_dict(golden_weights(fx), strict=False)
    m = m.cuda().train()
    B, n = args.mols, 60
    g = torch.Generator().manual_seed(1)
    pos, v = (3.0 * torch.randn(B * n, 3, generator=g)).cuda(), torch.randint(0, 15, (B * n,), generator=g).cuda()
    batch = torch.arange(B).repeat_interleave(n).cuda()
    shape = (0.07 * torch.randn(B, 32, 3, generator=g)).cuda()
    s = Sampler(m._engine(), pos, v, batch, shape, num_steps=1000, noise='philox', seed=3, keep_traj=False, use_graph=True, n_mols=B)
    s._step_body(0)
    s._capture(1)
    s.graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        s.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({'path': 'generic fp32 (hidden 256, k 48, 60 atoms)', 'molecules': B, 'ms_per_step': ms, 'mol_steps_per_s': B / (ms * 1e-3),
                      'finite': bool(torch.isfinite(s.pos).all())}))


if __name__ == '__main__':
    main()
