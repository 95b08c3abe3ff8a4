"""Small end-to-end run for compute-sanitizer (memcheck):  compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
from conftest import load_golden  # noqa: E402
from test_gpu_parity import build_model, batch_of  # noqa: E402
from shapemol_b200.engine import BatchDesc  # noqa: E402

fx = load_golden('forward_k32_train.pt')
g = torch.Generator().manual_seed(3)
for precision in ('bf16', 'bf16x3'):
    m = build_model(fx, precision)
    B = 333                                   # 2.6 node tiles per ... ragged, not a multiple of anything
    sizes = torch.randint(1, 28, (B,), generator=g).tolist()
    N = sum(sizes)
    pos, v = (2 * torch.randn(N, 3, generator=g)).cuda(), torch.randint(0, 15, (N,), generator=g).cuda()
    shape = (0.07 * torch.randn(B, 32, 3, generator=g)).cuda()
    m.smb_noise, m.smb_keep_traj = 'philox', True
    cloud = torch.randn(200, 3, generator=g, dtype=torch.float64).numpy()
    r = m.sample_diffusion(pos, v, batch_of(sizes), shape.view(-1, 3), num_steps=3, center_pos_mode='none',
                           use_pointcloud_data=(cloud, None, 0.3), grad_step=0)
    torch.cuda.synchronize()
    bd = BatchDesc(batch_of(sizes), B)
    ref = torch.randn(B * 7, 3, generator=g, dtype=torch.float64).cuda()
    rptr = torch.arange(0, B * 7 + 1, 7, dtype=torch.int32).cuda()
    t = m._engine().shape_tanimoto(bd, r['pos'].contiguous(), ref, rptr)
    torch.cuda.synchronize()
    print(precision, 'finite', bool(torch.isfinite(r['pos']).all()), 'tanimoto mean %.4f' % float(t.mean()))
print('sanitize run done')
