"""Point-cloud shape guidance through the C ABI against the reference fixtures and the oracle."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

from conftest import load_golden  # noqa: E402
from oracle import shapemol_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def engine_and_batch(n_atoms, sizes=None):
    from test_gpu_parity import build_model, batch_of
    from shapemol_b200.engine import BatchDesc
    fx = load_golden('forward_k8_eval.pt')
    m = build_model(fx, 'bf16x3', training=False)
    sizes = sizes or ([25] * (n_atoms // 25) + ([n_atoms % 25] if n_atoms % 25 else []))
    return m, m._engine(), BatchDesc(batch_of(sizes), len(sizes))


def test_guidance_matches_reference_fixture_bit_exact(cuda_lib):
    cases = torch.load(os.path.join(ROOT, 'tests', 'golden', 'guidance.pt'))
    for c in cases:
        N = c['pos'].shape[0]
        _, eng, bd = engine_and_batch(N)
        pos = c['pos'].cuda().clone()
        u = torch.nan_to_num(c['u'], nan=0.5).cuda().contiguous()
        eng.guidance(bd, pos, c['cloud'].cuda().contiguous(), c['radius'], ratio=0.2, u=u)
        torch.cuda.synchronize()
        assert torch.equal(pos.cpu(), c['out']), 'max diff %g' % float((pos.cpu() - c['out']).abs().max())


def test_guidance_ragged_clouds_time_gate_and_philox(cuda_lib):
    g = torch.Generator().manual_seed(5)
    sizes = [9, 27, 1, 14, 20]
    N, B = sum(sizes), len(sizes)
    _, eng, bd = engine_and_batch(N, sizes)
    counts = [60, 200, 5, 120, 2]                      # the last molecule's cloud is too small (< 3 points): untouched
    cloud = torch.randn(sum(counts), 3, generator=g, dtype=torch.float64)
    cptr = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32).cuda()
    pos0 = (2.5 * torch.randn(N, 3, generator=g)).float()
    t = torch.tensor([500, 100, 500, 301, 900], dtype=torch.int32).cuda()      # grad_step 300: molecule 1 is not guided
    u = torch.rand(5, N, generator=g, dtype=torch.float64)
    pos = pos0.cuda().clone()
    eng.guidance(bd, pos, cloud.cuda(), 0.4, t_i32=t, grad_step=300, u=u.cuda(), cloud_ptr=cptr)
    got = pos.cpu()
    ptr = [0]
    for s in sizes:
        ptr.append(ptr[-1] + s)
    cp = cptr.cpu().tolist()
    for m in range(B):
        a0, a1 = ptr[m], ptr[m + 1]
        if m == 1 or m == 4:
            assert torch.equal(got[a0:a1], pos0[a0:a1])
        else:
            exp = orc.pointcloud_guidance(pos0[a0:a1], cloud[cp[m]:cp[m + 1]], 0.4, u[:, a0:a1])
            assert torch.equal(got[a0:a1], exp)
    # in-kernel Philox: deterministic in (seed, atom, t), different seeds differ, every moved atom ends nearer its cloud
    outs = []
    for seed in (7, 7, 8):
        p = pos0.cuda().clone()
        eng.guidance(bd, p, cloud.cuda(), 0.4, t_i32=t, grad_step=300, cloud_ptr=cptr, seed=seed)
        outs.append(p.cpu())
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
    assert torch.isfinite(outs[0]).all()
    a0, a1 = ptr[0], ptr[1]
    d_before, _ = orc._three_nn(pos0[a0:a1].double(), cloud[cp[0]:cp[1]])
    d_after, _ = orc._three_nn(outs[0][a0:a1].double(), cloud[cp[0]:cp[1]])
    moved = (outs[0][a0:a1] != pos0[a0:a1]).any(1)
    assert moved.any() and bool((d_after.mean(1)[moved] < d_before.mean(1)[moved]).all())


def test_sampler_with_pointcloud_guidance_runs_through_dropin(cuda_lib):
    """sample_diffusion(use_pointcloud_data=(cloud, kdtree, radius), grad_step=...) as scripts/sample_diffusion.py:262-276."""
    from test_gpu_parity import build_model, batch_of
    fx = load_golden('forward_k8_eval.pt')
    m = build_model(fx, 'bf16x3', training=False)
    m.smb_noise, m.smb_seed, m.smb_keep_traj = 'philox', 3, True
    sizes = fx['sizes']
    batch = batch_of(sizes)
    g = torch.Generator().manual_seed(1)
    cloud = (0.5 * torch.randn(300, 3, generator=g, dtype=torch.float64)).numpy()
    T = m.num_timesteps
    kw = dict(init_ligand_pos=fx['pos'].cuda(), init_ligand_v=fx['v'].cuda(), batch_ligand=batch, ligand_shape=fx['shape'].view(-1, 3).cuda(),
              num_steps=6, center_pos_mode='none')
    plain = m.sample_diffusion(**kw)
    guided = m.sample_diffusion(use_pointcloud_data=(cloud, None, 0.2), grad_step=T - 4, **kw)     # steps t = T-1 .. T-3 are guided
    late = m.sample_diffusion(use_pointcloud_data=(cloud, None, 0.2), grad_step=T, **kw)           # t > T never holds: identical to plain
    assert torch.isfinite(guided['pos']).all() and len(guided['pos_traj']) == 6
    assert torch.equal(late['pos'], plain['pos'])
    assert not torch.equal(guided['pos'], plain['pos'])
    # the guided x0 prediction of the first step lies within / nearer the cloud than the unguided one
    c = torch.from_numpy(cloud)
    d_plain, _ = orc._three_nn(plain['pos_cond_traj'][0].double().cpu(), c)
    d_guided, _ = orc._three_nn(guided['pos_cond_traj'][0].double().cpu(), c)
    assert float(d_guided.mean()) < float(d_plain.mean())
