"""Golden vectors for the condition-shape front end from the UNMODIFIED reference functions
utils/shape.py:get_pointcloud_from_mol (:164-173) and get_pointAE_shape_emb (:240-284).

utils/shape.py cannot be imported here (rdkit / oddt / pytorch3d / trimesh are absent), so the two function definitions are
taken out of the reference file with `ast` at run time and executed unchanged in a namespace that provides numpy / torch and,
for get_pointAE_shape_emb, stand-ins for the two third-party-backed helpers it calls (get_mesh -> a prescribed (verts, faces),
get_pointcloud_from_mesh -> prescribed sampled points plus an object with pytorch3d's get_bounding_boxes()).  What is pinned is
therefore the reference's own arithmetic: the Gaussian clouds, the centring, the bounds frame, the batching and the encoder call.

Run in the build container only:   python tests/golden/make_frontend_golden.py   ->  tests/golden/frontend.pt
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('SHAPEMOL_REFERENCE', '/root/reference')


def reference_functions(names):
    src = open(os.path.join(REF, 'utils', 'shape.py')).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
    return out


class FakeMesh:
    def __init__(self, verts):
        self.verts = verts

    def get_bounding_boxes(self):     # pytorch3d: [1, 3, 2] (min | max per axis)
        return torch.stack([self.verts.min(0).values, self.verts.max(0).values], dim=1).unsqueeze(0)


class FakeAE:
    """encode(list of [b,1,P,3]) -> [B,4,3]: a deterministic function of the centred clouds (checks what reaches the encoder)."""

    def encode(self, batches):
        return torch.cat([torch.stack([b[:, 0].mean(1), b[:, 0].std(1), b[:, 0].amax(1), b[:, 0].amin(1)], dim=1) for b in batches], 0)


def main():
    fns = reference_functions(['get_pointcloud_from_mol', 'get_pointAE_shape_emb'])
    g = torch.Generator().manual_seed(7)
    out = {}
    # ---- get_pointcloud_from_mol ----
    ns = {'np': np}
    exec(fns['get_pointcloud_from_mol'], ns)
    poses = (3.0 * torch.randn(23, 3, generator=g)).double().numpy()
    np.random.seed(123)
    out['pc_from_mol'] = dict(poses=torch.from_numpy(poses), seed=123, N=20, points=torch.from_numpy(ns['get_pointcloud_from_mol'](poses)))
    # ---- get_pointAE_shape_emb (orchestration) ----
    n_mols, P = 7, 64
    verts = [4.0 * torch.randn(40 + 3 * i, 3, generator=g) + torch.tensor([1.0 * i, -2.0, 0.5]) for i in range(n_mols)]
    faces = [torch.randint(0, v.shape[0], (60, 3), generator=g) for v in verts]
    samples = [v[torch.randint(0, v.shape[0], (P,), generator=g)] + 0.01 * torch.randn(P, 3, generator=g) for v in verts]
    state = {'i': 0}

    def get_mesh(mol):
        return verts[mol].numpy(), faces[mol].numpy()

    def get_pointcloud_from_mesh(mesh, num_samples, return_mesh=False):
        i = state['i']
        state['i'] += 1
        return samples[i].unsqueeze(0), FakeMesh(torch.from_numpy(np.asarray(mesh[0])))

    ns = {'torch': torch, 'np': np, 'get_mesh': get_mesh, 'get_pointcloud_from_mesh': get_pointcloud_from_mesh}
    exec(fns['get_pointAE_shape_emb'], ns)
    zs, bounds, clouds, centers = ns['get_pointAE_shape_emb'](list(range(n_mols)), FakeAE(), P, None, shape_parallel=True, batch_size=3)
    out['shape_emb'] = dict(verts=verts, faces=faces, samples=samples, batch_size=3, zs=zs, bounds=bounds, clouds=clouds, centers=centers)
    torch.save(out, os.path.join(HERE, 'frontend.pt'))
    print('frontend.pt:', out['pc_from_mol']['points'].shape, zs.shape, bounds.shape, [c.shape for c in clouds], centers.shape)


if __name__ == '__main__':
    main()
