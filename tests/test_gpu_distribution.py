"""Third parity criterion of the north star: final distributions of a full reverse process are statistically
indistinguishable between the reference formulation (CPU oracle, torch noise) and the CUDA path (in-kernel Philox noise),
judged on pair distances, atom types and the reference's own alignment-free shape Tanimoto (get_ROCS).
A 50-step schedule keeps the CPU side at a few seconds; the weights are the synthetic fixture weights (the trained
checkpoint is not in the reference checkout), eval-mode BatchNorm so that molecules are independent samples."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

from conftest import load_golden, golden_weights, oracle_cfg  # noqa: E402
from oracle import shapemol_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
T = 50


def tanimoto_golden_and_ragged(eng, BatchDesc, batch_of):
    cases = torch.load(os.path.join(ROOT, 'tests', 'golden', 'rocs.pt'))
    cases = [c for c in cases if c['a'].shape[0] <= 60]
    sizes = [int(c['a'].shape[0]) for c in cases]
    pos = torch.cat([c['a'] for c in cases]).float().cuda()
    ref = torch.cat([c['b'] for c in cases]).cuda().contiguous()
    rptr = torch.tensor([0] + list(torch.tensor([int(c['b'].shape[0]) for c in cases]).cumsum(0)), dtype=torch.int32).cuda()
    bd = BatchDesc(batch_of(sizes), len(sizes))
    got = eng.shape_tanimoto(bd, pos, ref, rptr).cpu()
    # positions reach the kernel as float32 (the sampler's dtype): compare with the oracle on the same rounded inputs
    exp = torch.stack([orc.get_rocs(c['a'].float(), c['b']) for c in cases])
    assert float((got - exp).abs().max()) < 1e-12
    assert float((got - torch.stack([c['rocs'] for c in cases])).abs().max()) < 1e-6


def test_shape_tanimoto_kernel_matches_reference(cuda_lib):
    from test_gpu_parity import build_model, batch_of
    from shapemol_b200.engine import BatchDesc
    fx = load_golden('forward_k8_eval.pt')
    m = build_model(fx, 'bf16x3', training=False)
    tanimoto_golden_and_ragged(m._engine(), BatchDesc, batch_of)


def _population_stats(pos, v, mol_ptr, ref, ref_ptr):
    pos, v = pos.cpu(), v.cpu()
    d = []
    for m in range(len(mol_ptr) - 1):
        p = pos[int(mol_ptr[m]):int(mol_ptr[m + 1])].double()
        if p.shape[0] > 1:
            dm = (p[:, None] - p[None]).norm(dim=-1)
            d.append(dm[torch.triu(torch.ones_like(dm), 1) > 0])
    return torch.cat(d), torch.bincount(v.long(), minlength=15).double(), orc.get_rocs_batch(pos, mol_ptr, ref, ref_ptr)


def test_final_distributions_indistinguishable(cuda_lib):
    from scipy import stats
    from test_host_cpu import make_dropin
    from test_gpu_parity import batch_of
    from shapemol_b200.engine import BatchDesc
    fx = load_golden('forward_k32_eval.pt')
    sd = golden_weights(fx)
    g = torch.Generator().manual_seed(77)
    B = 64
    sizes = torch.randint(9, 28, (B,), generator=g).tolist()
    mol_ptr = orc.mol_ptr_from_sizes(sizes)
    N = int(mol_ptr[-1])
    shape = 0.07 * torch.randn(B, 32, 3, generator=g)
    # condition "reference molecules" for the Tanimoto statistic: 20 fixed centres per molecule
    ref = (1.5 * torch.randn(B * 20, 3, generator=g)).double()
    ref_ptr = torch.arange(0, B * 20 + 1, 20)

    def init(seed):
        gg = torch.Generator().manual_seed(seed)
        return torch.randn(N, 3, generator=gg), torch.randint(0, 15, (N,), generator=gg)

    # ---- population A: CPU oracle, torch noise ----
    cfg = dict(oracle_cfg(fx), num_diffusion_timesteps=T)
    tabs = orc.schedule_tables(T, cfg['schedule_pos'], cfg['schedule_v'])
    pos0, v0 = init(1)
    gn = torch.Generator().manual_seed(2)
    with torch.no_grad():
        pa, va, _ = orc.sample(sd, cfg, tabs, pos0, v0, mol_ptr, shape, T - 1, T,
                               lambda s: (torch.randn(N, 3, generator=gn), torch.rand(N, 15, generator=gn)), training=False)
    da, ha, ra = _population_stats(pa, va, mol_ptr, ref, ref_ptr)
    for precision in ('bf16x3', 'bf16'):
        # ---- population B: CUDA path, in-kernel Philox noise, different initial noise ----
        m, _ = make_dropin(knn=fx['k'], num_diffusion_timesteps=T)
        m.load_state_dict(sd, strict=False)
        m = m.cuda().eval()
        m.smb_precision, m.smb_noise, m.smb_seed, m.smb_keep_traj = precision, 'philox', 11, False
        pos1, v1 = init(3)
        r = m.sample_diffusion(pos1.cuda(), v1.cuda(), batch_of(sizes), shape.view(-1, 3).cuda(), num_steps=T, center_pos_mode='none')
        pb, vb = r['pos'], r['v']
        assert torch.isfinite(pb).all()

        db, hb, rb = _population_stats(pb, vb, mol_ptr, ref, ref_ptr)
        # the device Tanimoto kernel gives the same statistic as the oracle on the device population
        bd = BatchDesc(batch_of(sizes), B)
        rb_dev = m._engine().shape_tanimoto(bd, pb.contiguous(), ref.cuda().contiguous(), ref_ptr.to(torch.int32).cuda()).cpu()
        assert float((rb_dev - rb).abs().max()) < 1e-9
        p_dist = stats.ks_2samp(da.numpy(), db.numpy()).pvalue
        p_rocs = stats.ks_2samp(ra.numpy(), rb.numpy()).pvalue
        keep = (ha + hb) > 0
        p_type = stats.chi2_contingency(torch.stack([ha[keep], hb[keep]]).numpy())[1]
        # sanity: the statistics can tell populations apart -- the initial noise is NOT distributed like the final samples
        d0, _, r0 = _population_stats(pos0, v0, mol_ptr, ref, ref_ptr)
        p_null = stats.ks_2samp(da.numpy(), d0.numpy()).pvalue
        print('%s: KS p(pair distances) %.3f  KS p(shape Tanimoto) %.3f  chi2 p(atom types) %.3f  | init-vs-final p %.2e'
              % (precision, p_dist, p_rocs, p_type, p_null))
        assert p_null < 1e-6
        assert p_dist > 0.01 and p_rocs > 0.01 and p_type > 0.01


def _molecule_stats(pos, v, mol_ptr, ref, ref_ptr):
    """Per-molecule statistics (independent samples, unlike the pooled pair distances of one coupled batch): radius of gyration,
    mean pair distance, mean nearest-neighbour distance, shape Tanimoto; plus the atom-type histogram."""
    pos, v = pos.cpu().double(), v.cpu()
    rg, mp, nn = [], [], []
    for m in range(len(mol_ptr) - 1):
        p = pos[int(mol_ptr[m]):int(mol_ptr[m + 1])]
        dm = (p[:, None] - p[None]).norm(dim=-1)
        rg.append((p - p.mean(0)).pow(2).sum(1).mean().sqrt())
        mp.append(dm[torch.triu(torch.ones_like(dm), 1) > 0].mean())
        nn.append((dm + 1e9 * torch.eye(p.shape[0], dtype=dm.dtype)).min(1).values.mean())
    return dict(rg=torch.stack(rg), pair=torch.stack(mp), nn=torch.stack(nn), rocs=orc.get_rocs_batch(pos.float(), mol_ptr, ref, ref_ptr),
                types=torch.bincount(v.long(), minlength=15).double())


def test_full_1000_step_population_vs_unmodified_reference(cuda_lib):
    """The process the reference actually runs: 1000 reverse steps with train-mode BatchNorm, final state compared with
    tests/golden/population.pt -- eight independent runs of the UNMODIFIED reference (tests/golden/make_population_golden.py).
    Batch statistics couple the 48 molecules of a run (the run-level mean radius of gyration scatters by 0.1 .. 0.25 A between
    runs, against 0.07 A if molecules were independent), so the RUN is the sampling unit: eight CUDA runs per arithmetic mode
    (in-kernel Philox noise) against the eight reference runs, Mann-Whitney on the run-level means of every statistic, and the
    pooled atom-type histogram.  Not distinguishable at the 1 % level = pass."""
    from scipy import stats
    from test_host_cpu import make_dropin
    from test_gpu_parity import batch_of
    import synth
    from conftest import manifest_shapes
    fx = torch.load(os.path.join(ROOT, 'tests', 'golden', 'population.pt'))
    sizes, shape, steps = fx['sizes'], fx['shape'], fx['steps']
    B, R = len(sizes), len(fx['runs'])
    assert R >= 8
    mol_ptr = orc.mol_ptr_from_sizes(sizes)
    N = int(mol_ptr[-1])
    g = torch.Generator().manual_seed(5)
    ref = (1.5 * torch.randn(B * 20, 3, generator=g)).double()
    ref_ptr = torch.arange(0, B * 20 + 1, 20)
    keys = ('rg', 'pair', 'nn', 'rocs')

    def run_level(pops):
        return {k: torch.stack([p[k].mean() for p in pops]) for k in keys}, sum(p['types'] for p in pops)
    assert all(bool(torch.isfinite(r['pos']).all()) for r in fx['runs'])
    ref_runs, ref_types = run_level([_molecule_stats(r['pos'], r['v'], mol_ptr, ref, ref_ptr) for r in fx['runs']])
    # the statistics can tell populations apart: the initial noise is not distributed like the final samples
    init_runs, _ = run_level([_molecule_stats(r['pos0'], r['v0'], mol_ptr, ref, ref_ptr) for r in fx['runs']])
    assert stats.mannwhitneyu(ref_runs['rg'].numpy(), init_runs['rg'].numpy()).pvalue < 1e-3
    print('reference runs: ' + '  '.join('%s %.3f +- %.3f' % (k, float(ref_runs[k].mean()), float(ref_runs[k].std())) for k in keys))
    sd = synth.synth_state_dict(manifest_shapes(), fx['seed'])
    for precision in ('bf16x3', 'bf16'):
        pops = []
        for run in range(R):
            m, _ = make_dropin(knn=fx['k'], num_diffusion_timesteps=steps)
            m.load_state_dict(sd, strict=False)
            m = m.cuda().train()
            m.smb_precision, m.smb_noise, m.smb_seed, m.smb_keep_traj = precision, 'philox', 100 + run, False
            gg = torch.Generator().manual_seed(700 + run)
            pos1, v1 = torch.randn(N, 3, generator=gg), torch.randint(0, 15, (N,), generator=gg)
            r = m.sample_diffusion(pos1.cuda(), v1.cuda(), batch_of(sizes), shape.view(-1, 3).cuda(), num_steps=steps, center_pos_mode='none')
            assert torch.isfinite(r['pos']).all()
            pops.append(_molecule_stats(r['pos'], r['v'], mol_ptr, ref, ref_ptr))
        mine, my_types = run_level(pops)
        pp = {k: stats.mannwhitneyu(mine[k].numpy(), ref_runs[k].numpy(), alternative='two-sided').pvalue for k in keys}
        keep = (my_types + ref_types) > 0
        pp['types'] = stats.chi2_contingency(torch.stack([my_types[keep], ref_types[keep]]).numpy())[1]
        print('%s runs:      ' % precision + '  '.join('%s %.3f +- %.3f' % (k, float(mine[k].mean()), float(mine[k].std())) for k in keys)
              + ' | p: ' + '  '.join('%s %.3f' % kv for kv in pp.items()))
        assert min(pp.values()) > 0.01, (precision, pp)


def test_stability_kernel_matches_reference(cuda_lib):
    from test_gpu_parity import build_model, batch_of
    from shapemol_b200.engine import BatchDesc
    from shapemol_b200 import chem_tables as ct
    cases = torch.load(os.path.join(ROOT, 'tests', 'golden', 'stability.pt'))
    fx = load_golden('forward_k8_eval.pt')
    eng = build_model(fx, 'bf16x3', training=False)._engine()
    for hs in (False, True):
        cs = [c for c in cases if c['hs'] == hs]
        sizes = [int(c['pos'].shape[0]) for c in cs]
        bd = BatchDesc(batch_of(sizes), len(sizes))
        pos = torch.cat([c['pos'] for c in cs]).cuda()
        z = torch.cat([c['z'] for c in cs]).cuda()
        mol_stable, st_atoms, nr = eng.check_stability(bd, pos, z, hs=hs)
        assert torch.equal(nr.cpu().long(), torch.cat([c['nr_bonds'] for c in cs]))
        assert st_atoms.cpu().tolist() == [c['nr_stable'] for c in cs]
        assert mol_stable.cpu().tolist() == [c['stable'] for c in cs]
