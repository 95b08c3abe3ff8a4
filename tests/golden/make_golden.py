"""Generates the committed golden fixtures by running the UNMODIFIED reference modules
(/root/reference/models/*.py) on CPU through the shims in tests/golden/_shims.

Run in the build container only:   python tests/golden/make_golden.py
Outputs: tests/golden/{manifest.json, schedules.pt, forward_*.pt, knn_cases.pt, trajectory.pt, encoder.pt}
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
import synth  # noqa: E402

torch.set_num_threads(8)


def build_model(msm, ED, seed, training, **over):
    cfg = ref_loader.model_config(ED, **over)
    torch.manual_seed(seed)
    m = msm.ScorePosNet3D(cfg, ligand_atom_feature_dim=15)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(synth.is_non_synth(k) for k in missing), missing
    m.train(training)
    return m, shapes


def batch_of(sizes):
    return torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))


def forward_case(msm, ED, name, seed, sizes, k, training, t_vals, hidden=128, heads=16):
    over = dict(knn=k)
    if hidden != 128:
        over.update(hidden_dim=hidden, n_heads=heads)
    m, shapes = build_model(msm, ED, seed, training, **over)
    g = torch.Generator().manual_seed(seed + 17)
    pos = synth.molecule_like_positions(sizes, seed + 3)
    N = pos.shape[0]
    v = torch.randint(0, 15, (N,), generator=g)
    shape = 0.07 * torch.randn(len(sizes), 32, 3, generator=g)
    t = torch.tensor(t_vals, dtype=torch.long)
    batch = batch_of(sizes)
    captured = {}
    orig = m.refine_net._connect_edge

    def spy(*a, **kw):
        e = orig(*a, **kw)
        captured['edge_index'] = e.clone()
        return e
    m.refine_net._connect_edge = spy
    with torch.no_grad():
        out = m(pos, v, batch, shape, time_step=t)
    fx = dict(seed=seed, sizes=sizes, k=k, training=training, hidden=hidden, heads=heads,
              pos=pos, v=v, shape=shape, t=t, edge_index=captured['edge_index'].int(),
              pred_pos=out['pred_ligand_pos'], pred_h=out['pred_ligand_h'], pred_v=out['pred_ligand_v'])
    if hidden != 128:
        fx['shapes'] = {k: list(v) for k, v in shapes.items()}
    if training:   # running-stat side effect of train-mode BN (layer 0 and 7)
        sd = m.state_dict()
        for l in (0, 7):
            p = 'refine_net.base_block.%d.h2x_layers.0.shape_linear.batchnorm.bn.' % l
            fx['bn%d_running_mean' % l] = sd[p + 'running_mean'].clone()
            fx['bn%d_running_var' % l] = sd[p + 'running_var'].clone()
    torch.save(fx, os.path.join(HERE, 'forward_%s.pt' % name))
    print(name, 'N=%d E=%d' % (N, captured['edge_index'].shape[1]))


def knn_cases():
    """Tie / degenerate kNN cases through the same knn_graph the reference calls (shim restatement
    of torch_cluster semantics) -- recorded so the CUDA kernel and the oracle are checked against a
    committed table, not against each other only."""
    from torch_geometric.nn import knn_graph
    cases = {}
    g = torch.Generator().manual_seed(5)
    lattice = torch.stack(torch.meshgrid(torch.arange(3.), torch.arange(3.), torch.arange(3.), indexing='ij'), -1).view(-1, 3)
    dup = torch.randn(12, 3, generator=g)
    dup[5] = dup[2]
    dup[9] = dup[2]
    line = torch.stack([torch.arange(10.), torch.zeros(10), torch.zeros(10)], 1)
    many_dup = torch.zeros(14, 3)
    many_dup[10:] = torch.randn(4, 3, generator=g)
    big = torch.randn(60, 3, generator=g) * 3
    for name, x, k in [('lattice_k8', lattice, 8), ('lattice_k32', lattice, 32), ('dup_k4', dup, 4),
                       ('line_k3', line, 3), ('many_dup_k8', many_dup, 8), ('big_k48', big, 48),
                       ('single', torch.randn(1, 3, generator=g), 8), ('pair', torch.randn(2, 3, generator=g), 8)]:
        e = knn_graph(x, k=k, batch=torch.zeros(x.shape[0], dtype=torch.long), flow='source_to_target')
        cases[name] = dict(x=x, k=k, edge_index=e.int())
    torch.save(cases, os.path.join(HERE, 'knn_cases.pt'))


def trajectory_case(msm, ED):
    seed, sizes, k, steps = 11, [13, 27, 21, 9, 24, 18, 26, 22], 32, 12
    m, _ = build_model(msm, ED, seed, True, knn=k)
    g = torch.Generator().manual_seed(99)
    N = sum(sizes)
    pos0 = torch.randn(N, 3, generator=g)
    v0 = torch.randint(0, 15, (N,), generator=g)
    shape = 0.07 * torch.randn(len(sizes), 32, 3, generator=g)
    noise = [(torch.randn(N, 3, generator=g), torch.rand(N, 15, generator=g)) for _ in range(steps)]
    it = iter(noise)
    # inject the noise in the reference's draw order: randn_like(pos) then rand_like(logits)
    orig_randn_like, orig_rand_like = torch.randn_like, torch.rand_like
    state = {}

    def randn_like(x, *a, **kw):
        state['cur'] = next(it)
        return state['cur'][0].to(x)

    def rand_like(x, *a, **kw):
        return state['cur'][1].to(x)
    torch.randn_like, torch.rand_like = randn_like, rand_like
    try:
        r = m.sample_diffusion(init_ligand_pos=pos0, init_ligand_v=v0, batch_ligand=batch_of(sizes),
                               ligand_shape=shape.view(-1, 3), num_steps=steps, center_pos_mode='none')
    finally:
        torch.randn_like, torch.rand_like = orig_randn_like, orig_rand_like
    fx = dict(seed=seed, sizes=sizes, k=k, steps=steps, pos0=pos0, v0=v0, shape=shape,
              noise_pos=torch.stack([n[0] for n in noise]), noise_u=torch.stack([n[1] for n in noise]),
              pos=r['pos'], v=r['v'], pos_traj=torch.stack(r['pos_traj']), v_traj=torch.stack(r['v_traj']),
              pos_cond_traj=torch.stack(r['pos_cond_traj']), v_cond_traj=torch.stack(r['v_cond_traj']),
              v0_traj=torch.stack(r['v0_traj']), vt_traj=torch.stack(r['vt_traj']))
    torch.save(fx, os.path.join(HERE, 'trajectory.pt'))
    print('trajectory', fx['pos_traj'].shape)


def encoder_case(spm, ED):
    ck = torch.load(os.path.join(ref_loader.REF_ROOT, 'trained_models/se_model.pt'), map_location='cpu',
                    weights_only=False)
    torch.manual_seed(3)
    ae = spm.PointCloud_AE(ck['config'].model)
    ae.load_state_dict(ck['model'], strict=True)
    enc = ae.encoder
    shapes = {}
    for i, b in enumerate(enc.blocks):
        for k, v in b.state_dict().items():
            shapes['blocks.%d.%s' % (i, k)] = tuple(v.shape)
    bw = synth.synth_state_dict(shapes, 3, skip_non_synth=False)
    for i, b in enumerate(enc.blocks):
        b.load_state_dict({k[len('blocks.%d.' % i):]: v for k, v in bw.items() if k.startswith('blocks.%d.' % i)})
    trained = {k[len('encoder.'):]: v.clone() for k, v in ck['model'].items() if k.startswith('encoder.')}
    g = torch.Generator().manual_seed(8)
    fx = dict(trained=trained, block_shapes=shapes, num_k=int(ck['config'].model.num_k))
    for name, B, P in [('small', 3, 128), ('p512', 2, 512)]:
        clouds = torch.randn(B, 1, P, 3, generator=g) * torch.tensor([3.0, 2.0, 1.5])
        clouds = clouds - clouds.mean(2, keepdim=True)
        for mode in ('train', 'eval'):
            ae.load_state_dict(ck['model'], strict=True)   # undo the running-stat side effect of a train pass
            enc.conv_pos.train(mode == 'train')
            enc.conv_c.train(mode == 'train')
            with torch.no_grad():
                lat = enc(clouds)
            fx['%s_%s_latent' % (name, mode)] = lat
        fx['%s_clouds' % name] = clouds
    torch.save(fx, os.path.join(HERE, 'encoder.pt'))
    print('encoder', {k: tuple(v.shape) for k, v in fx.items() if torch.is_tensor(v)})


def main():
    msm, spm, ED = ref_loader.load()
    m, shapes = build_model(msm, ED, 1, True, knn=32)
    sd = m.state_dict()
    with open(os.path.join(HERE, 'manifest.json'), 'w') as f:
        json.dump({k: [list(v), str(sd[k].dtype)] for k, v in shapes.items()}, f, indent=0)
    torch.save({k: sd[k].clone() for k in synth.SCHEDULE_KEYS}, os.path.join(HERE, 'schedules.pt'))
    if '--encoder-only' in sys.argv:
        return encoder_case(spm, ED)
    forward_case(msm, ED, 'k32_train', 21, [27, 13, 19, 9, 24, 22], 32, True, [999, 500, 1, 0, 250, 750])
    forward_case(msm, ED, 'k32_eval', 22, [27, 27, 18, 26], 32, False, [400, 400, 400, 400])
    forward_case(msm, ED, 'k8_train', 23, [20, 27, 15, 23, 9], 8, True, [600, 30, 999, 0, 123])
    forward_case(msm, ED, 'k8_eval', 24, [21, 25, 12], 8, False, [77, 500, 900])
    forward_case(msm, ED, 'k48_h256_train', 25, [60, 52], 48, True, [300, 800], hidden=256, heads=16)
    forward_case(msm, ED, 'tiny_train', 26, [1, 2, 3, 5], 32, True, [10, 20, 30, 40])
    knn_cases()
    trajectory_case(msm, ED)
    encoder_case(spm, ED)


if __name__ == '__main__':
    main()
