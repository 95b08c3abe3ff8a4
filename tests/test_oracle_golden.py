"""Pins the oracle (oracle/shapemol_oracle.py) against fixtures produced by the UNMODIFIED
reference modules (tests/golden/make_golden.py).  CPU only."""
import os
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_weights, oracle_cfg
from oracle import shapemol_oracle as orc

FWD = ['k32_train', 'k32_eval', 'k8_train', 'k8_eval', 'k48_h256_train', 'tiny_train']


def test_schedule_tables_bit_exact():
    ref = load_golden('schedules.pt')
    tabs = orc.schedule_tables(1000, orc.DEFAULT_CFG['schedule_pos'], orc.DEFAULT_CFG['schedule_v'])
    for k, v in tabs.items():
        assert torch.equal(v, ref[k]), k


def test_knn_tie_cases_bit_exact():
    cases = load_golden('knn_cases.pt')
    for name, c in cases.items():
        n = c['x'].shape[0]
        e = orc.knn_edges(c['x'], torch.tensor([0, n]), c['k'])
        assert torch.equal(e.int(), c['edge_index']), name


@pytest.mark.parametrize('name', FWD)
def test_forward_matches_reference(name):
    fx = load_golden('forward_%s.pt' % name)
    sd = golden_weights(fx)
    cfg = oracle_cfg(fx)
    mol_ptr = orc.mol_ptr_from_sizes(fx['sizes'])
    nbr, stats = [], []
    with torch.no_grad():
        x, h, logits = orc.forward(sd, cfg, fx['pos'], fx['v'], mol_ptr, fx['shape'], fx['t'],
                                   training=fx['training'], bn_stats_out=stats, nbr_out=nbr)
    assert torch.equal(nbr[0].int(), fx['edge_index'])
    # same op order as the reference => agreement at fp32 round-off level
    assert torch.allclose(x, fx['pred_pos'], rtol=1e-5, atol=2e-5)
    assert torch.allclose(h, fx['pred_h'], rtol=1e-5, atol=2e-5)
    assert torch.allclose(logits, fx['pred_v'], rtol=1e-5, atol=2e-5)
    if fx['training']:
        for l in (0, 7):
            mean, var_unb = stats[l]
            rm = 0.9 * sd['refine_net.base_block.%d.h2x_layers.0.shape_linear.batchnorm.bn.running_mean' % l] + 0.1 * mean
            rv = 0.9 * sd['refine_net.base_block.%d.h2x_layers.0.shape_linear.batchnorm.bn.running_var' % l] + 0.1 * var_unb
            assert torch.allclose(rm, fx['bn%d_running_mean' % l], rtol=1e-5, atol=1e-6)
            assert torch.allclose(rv, fx['bn%d_running_var' % l], rtol=1e-4, atol=1e-6)


def test_trajectory_matches_reference():
    fx = load_golden('trajectory.pt')
    sd = golden_weights(fx)
    cfg = oracle_cfg(fx)
    tabs = orc.schedule_tables(1000, cfg['schedule_pos'], cfg['schedule_v'])
    mol_ptr = orc.mol_ptr_from_sizes(fx['sizes'])
    with torch.no_grad():
        pos, v, traj = orc.sample(sd, cfg, tabs, fx['pos0'], fx['v0'], mol_ptr, fx['shape'], 999, fx['steps'],
                                  lambda s: (fx['noise_pos'][s], fx['noise_u'][s]), training=True, keep_traj=True)
    for s, (x0, logits, p, vv, lv0, post) in enumerate(traj):
        assert torch.allclose(x0, fx['pos_cond_traj'][s], rtol=1e-4, atol=1e-4), s
        assert torch.allclose(logits, fx['v_cond_traj'][s], rtol=1e-4, atol=1e-4), s
        assert torch.allclose(p, fx['pos_traj'][s], rtol=1e-4, atol=1e-4), s
        assert torch.equal(vv, fx['v_traj'][s]), s
        assert torch.allclose(lv0, fx['v0_traj'][s], rtol=1e-4, atol=1e-4), s
        assert torch.allclose(post, fx['vt_traj'][s], rtol=1e-4, atol=2e-4), s
    assert torch.equal(v, fx['v'])


@pytest.mark.parametrize('name', ['small', 'p512'])
@pytest.mark.parametrize('mode', ['train', 'eval'])
def test_encoder_matches_reference(name, mode):
    import synth
    fx = load_golden('encoder.pt')
    w = dict(fx['trained'])
    w.update(synth.synth_state_dict(fx['block_shapes'], 3, skip_non_synth=False))
    with torch.no_grad():
        lat = orc.encoder_forward(w, fx['%s_clouds' % name], k=fx['num_k'], training=(mode == 'train'))
    assert torch.allclose(lat, fx['%s_%s_latent' % (name, mode)], rtol=1e-4, atol=1e-5)


def test_pointcloud_guidance_oracle_matches_reference_bit_for_bit():
    """tests/golden/guidance.pt: outputs of the unmodified pointcloud_shape_guidance (models/molopt_score_model.py:699-740)
    with sklearn's KDTree and recorded numpy draws (make_guidance_golden.py)."""
    cases = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'guidance.pt'))
    assert len(cases) == 4
    for c in cases:
        got = orc.pointcloud_guidance(c['pos'], c['cloud'], c['radius'], torch.nan_to_num(c['u'], nan=0.5))
        assert torch.equal(got, c['out'])
        # atoms that start within the radius are never touched; moved atoms end closer to the cloud's 3-NN centroid
        d, _ = orc._three_nn(c['pos'].double(), c['cloud'])
        near = d.mean(1) <= c['radius']
        assert torch.equal(got[near], c['pos'][near])


def test_shape_tanimoto_oracle_matches_reference():
    """tests/golden/rocs.pt: outputs of the unmodified get_ROCS (utils/evaluation/shaep_utils.py:59-83)."""
    cases = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rocs.pt'))
    for c in cases:
        assert abs(float(orc.get_rocs(c['a'], c['b'])) - float(c['rocs'])) < 1e-12
    assert abs(float(orc.get_rocs(cases[-1]['a'], cases[-1]['b'])) - 1.0) < 1e-12      # identical centre sets


def test_stability_oracle_and_tables_match_reference():
    """tests/golden/stability.pt: outputs of the unmodified check_stability (utils/evaluation/analyze.py:264-297)."""
    from shapemol_b200 import chem_tables as ct
    cases = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stability.pt'))
    thr, allowed = ct.thresholds(), torch.tensor(ct.ALLOWED_BONDS, dtype=torch.int32)
    assert thr.shape == (3, 10, 10) and bool((thr == thr.transpose(1, 2)).all())
    for c in cases:
        stable, nr_stable, n, nr = orc.check_stability(c['pos'], ct.element_index(c['z']), thr, allowed, hs=c['hs'])
        assert (stable, nr_stable, n) == (c['stable'], c['nr_stable'], c['pos'].shape[0]) and torch.equal(nr, c['nr_bonds'])
    with pytest.raises(KeyError):
        ct.element_index(torch.tensor([6, 5]))      # boron is not in the reference's tables
