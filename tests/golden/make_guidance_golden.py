"""Golden vectors for the point-cloud shape guidance from the UNMODIFIED reference function
models/molopt_score_model.py:699-740 (pointcloud_shape_guidance), run on CPU with sklearn's KDTree as the reference
script builds it (scripts/sample_diffusion.py:237-241).  numpy's global RNG is seeded and its draws are recorded, then
laid out densely as u[iteration][atom] (the function draws one scalar per still-far atom, ascending atom order).

Run in the build container only:   python tests/golden/make_guidance_golden.py   ->  tests/golden/guidance.pt
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_loader  # noqa: E402


def one_case(fn, seed, n_ref, n_atoms, spread, radius):
    from sklearn.neighbors import KDTree
    rng = np.random.RandomState(seed)
    ref_atoms = rng.randn(n_ref, 3) * 2.0
    # utils/shape.py:164-173 get_pointcloud_from_mol: 20 points per atom from N(atom, 1/(12*1.7) I)
    var = 1. / (12. * 1.7)
    cloud = np.concatenate([rng.multivariate_normal(ref_atoms[i], np.eye(3) * var, size=20) for i in range(n_ref)], axis=0)
    pos = (ref_atoms[rng.randint(0, n_ref, n_atoms)] + rng.randn(n_atoms, 3) * spread).astype(np.float32)
    draws = []
    orig_random = np.random.random
    np.random.seed(seed + 1000)

    def rec(n):
        r = orig_random(n)
        draws.append(r.copy())
        return r
    np.random.random = rec
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        inp = torch.from_numpy(pos.copy())
        out = fn((cloud, KDTree(cloud), radius), inp)
    finally:
        np.random.random = orig_random
        torch.Tensor.cuda = orig_cuda
    return dict(pos=torch.from_numpy(pos), cloud=torch.from_numpy(cloud), radius=float(radius), out=out.clone(), draws=draws)


def densify(case):
    """Replays the far sets with the oracle to place draw i of iteration j at u[j][atom]."""
    from oracle import shapemol_oracle as orc
    N = case['pos'].shape[0]
    u = torch.full((5, N), float('nan'), dtype=torch.float64)
    cloud = case['cloud']
    p = case['pos'].to(torch.float64)
    dists, idx = orc._three_nn(p, cloud)
    far = ((dists[:, 0] + dists[:, 1]) + dists[:, 2]) / 3.0 > case['radius']
    atoms = torch.nonzero(far).flatten()
    pts, nn = p[atoms], idx[atoms]
    j = 0
    while atoms.numel() > 0 and j < 5:
        r = torch.from_numpy(case['draws'][j])
        assert r.numel() == atoms.numel(), (j, r.numel(), atoms.numel())
        u[j, atoms] = r
        c = cloud[nn]
        near = ((c[:, 0] + c[:, 1]) + c[:, 2]) / 3.0
        new = pts - (r * (0.8 - 0.2) + 0.2)[:, None] * (pts - near)
        dists, idx = orc._three_nn(new, cloud)
        inside = ((dists[:, 0] + dists[:, 1]) + dists[:, 2]) / 3.0 < case['radius']
        atoms, pts, nn = atoms[~inside], new[~inside], idx[~inside]
        j += 1
    assert j == len(case['draws'])
    return u


def main():
    msm, _, _ = ref_loader.load()
    fn = msm.pointcloud_shape_guidance
    cases = []
    for seed, n_ref, n_atoms, spread, radius in ((1, 22, 200, 0.6, 0.2), (2, 13, 64, 1.5, 0.2), (3, 26, 500, 0.3, 0.35), (4, 9, 27, 3.0, 0.2)):
        c = one_case(fn, seed, n_ref, n_atoms, spread, radius)
        c['u'] = densify(c)
        from oracle import shapemol_oracle as orc
        got = orc.pointcloud_guidance(c['pos'], c['cloud'], c['radius'], torch.nan_to_num(c['u'], nan=0.5))
        assert torch.equal(got, c['out']), 'oracle does not reproduce the reference bit for bit (case %d)' % seed
        moved = int((c['out'] != c['pos']).any(1).sum())
        print('case seed %d: %d atoms, %d cloud points, %d moved, iterations %d' % (seed, n_atoms, c['cloud'].shape[0], moved, len(c['draws'])))
        del c['draws']
        cases.append(c)
    torch.save(cases, os.path.join(HERE, 'guidance.pt'))


if __name__ == '__main__':
    main()
