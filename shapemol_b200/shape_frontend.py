"""Condition-shape front end: from molecules / surfaces to the batch of shape latents the denoiser conditions on.

Mirrors the point-cloud branch of the reference's utils/shape.py:

    reference                                    here
    ------------------------------------------   ---------------------------------------------------------------
    get_pointcloud_from_mol        (:164-173)    get_pointcloud_from_mol      atom-centred Gaussians (guidance clouds)
    get_pointcloud_from_mesh       (:175-184)    sample_points_from_mesh      area-weighted surface sampling of a triangle mesh
    get_mesh  (ODDT marching cubes, :153-162)    sample_molecular_surface     MESH-FREE stand-in: points on the solvent-excluded
                                                                              contact surface of the van-der-Waals spheres
    get_pointAE_shape_emb          (:240-284)    get_pointAE_shape_emb        centre, bound, batch, encode (CUDA VN-DGCNN encoder)

Parity status.  get_pointcloud_from_mol and the centring / bounds / batching of get_pointAE_shape_emb are pinned against the
unmodified reference functions (tests/golden/make_frontend_golden.py -> tests/golden/frontend.pt).  The reference's surface
itself comes from third-party code that is absent here (oddt.surface.generate_surface_marching_cubes, pytorch3d
sample_points_from_meshes, pinned only by prose in ReadMe.md): sample_points_from_mesh restates pytorch3d's published algorithm
(multinomial over face areas, then u, v ~ U[0,1): w0 = 1 - sqrt(u), w1 = sqrt(u)(1 - v), w2 = sqrt(u) v) and
sample_molecular_surface is an approximation of the ODDT surface -- both are PARITY UNPINNED and tested through properties only.

Everything runs on the tensors' device (torch ops are plumbing here; the encoder is the CUDA library).
"""
import math

import numpy as np
import torch

# van der Waals radii, reference utils/shape.py:27-38 (constants)
ATOM_RADIUS = {'C': 1.908, 'F': 1.75, 'Cl': 1.948, 'Br': 2.22, 'I': 2.35, 'N': 1.824, 'O': 1.6612, 'P': 2.1, 'S': 2.0, 'Si': 2.2}
ATOMIC_NUMBER = {'C': 6, 'F': 9, 'Cl': 17, 'Br': 35, 'I': 53, 'N': 7, 'O': 8, 'P': 15, 'S': 16, 'Si': 14}
_RADIUS_BY_Z = {z: ATOM_RADIUS[s] for s, z in ATOMIC_NUMBER.items()}


def get_pointcloud_from_mol(poses, confId=-1, N=20, var=1. / (12. * 1.7), rng=None):
    """Atom-centred isotropic Gaussians, N points per atom -- reference utils/shape.py:164-173.

    The reference draws `np.random.multivariate_normal(poses[i], var * I, size=N)` per atom from NumPy's global generator.  For
    an isotropic covariance that is  poses[i] + sqrt(var) * standard_normal((N, 3)),  so with `rng = np.random` (default) or a
    `np.random.RandomState(seed)` this function consumes the generator exactly like the reference and returns the same float64
    [n_atoms * N, 3] array (to rounding of the reference's SVD factor, ~1e-16)."""
    rng = np.random if rng is None else rng
    poses = np.asarray(poses, dtype=np.float64)
    s = math.sqrt(var)
    out = np.empty((poses.shape[0] * N, 3), dtype=np.float64)
    for i in range(poses.shape[0]):
        out[i * N:(i + 1) * N] = poses[i] + s * rng.standard_normal((N, 3))
    return out


def sample_points_from_mesh(verts, faces, num_samples, generator=None):
    """Uniform-by-area samples of a triangle mesh: verts [V,3] float, faces [F,3] int -> [num_samples, 3] on verts.device.
    Restates pytorch3d.ops.sample_points_from_meshes as the reference calls it (utils/shape.py:175-184); parity unpinned."""
    verts = torch.as_tensor(verts, dtype=torch.float32)
    faces = torch.as_tensor(faces).long().to(verts.device)
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    areas = 0.5 * torch.linalg.cross(v1 - v0, v2 - v0).norm(dim=1)
    idx = torch.multinomial(areas.clamp_min(1e-12), num_samples, replacement=True, generator=generator)
    uv = torch.rand(2, num_samples, device=verts.device, generator=generator)
    su = uv[0].sqrt()
    w0, w1, w2 = 1.0 - su, su * (1.0 - uv[1]), su * uv[1]
    return w0[:, None] * v0[idx] + w1[:, None] * v1[idx] + w2[:, None] * v2[idx]


def mesh_bounds(verts):
    """[2,3] (min | max), the layout of pytorch3d Meshes.get_bounding_boxes().squeeze(0).transpose(1, 0)."""
    verts = torch.as_tensor(verts, dtype=torch.float32)
    return torch.stack([verts.min(0).values, verts.max(0).values], 0)


def sample_molecular_surface(coords, atomic_numbers, num_samples, probe_radius=1.4, generator=None, oversample=6):
    """Mesh-free stand-in for get_mesh (ODDT marching cubes, probe 1.4 A) + surface sampling: uniform-by-area points on the
    solvent-ACCESSIBLE surface (union of spheres of radius r_vdw + probe), pulled back by the probe radius along the sphere
    normal -- the contact part of the solvent-excluded surface.  coords [n,3], atomic_numbers [n] -> ([num_samples,3], bounds [2,3]).
    Approximation (re-entrant patches are not reproduced); parity unpinned."""
    coords = torch.as_tensor(coords, dtype=torch.float32)
    dev = coords.device
    r = torch.tensor([_RADIUS_BY_Z.get(int(z), 1.9) for z in atomic_numbers], dtype=torch.float32, device=dev)
    R = r + probe_radius
    n = coords.shape[0]
    total = num_samples * oversample
    while True:
        # candidates: atom chosen by expanded-sphere area, uniform direction on the sphere
        atom = torch.multinomial(R * R, total, replacement=True, generator=generator)
        d = torch.randn(total, 3, device=dev, generator=generator)
        d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-12)
        p = coords[atom] + R[atom, None] * d
        # keep the points outside every other expanded sphere
        dist = torch.cdist(p, coords)                                   # [total, n]
        dist[torch.arange(total, device=dev), atom] = float('inf')
        keep = (dist >= R[None, :] - 1e-4).all(dim=1)
        p, atom, d = p[keep], atom[keep], d[keep]
        if p.shape[0] >= num_samples or n == 0:
            break
        total *= 2
    sel = torch.randperm(p.shape[0], device=dev, generator=generator)[:num_samples]
    pts = p[sel] - probe_radius * d[sel]
    lo = (coords - r[:, None]).min(0).values
    hi = (coords + r[:, None]).max(0).values
    return pts, torch.stack([lo, hi], 0)


def get_pointAE_shape_emb(surfaces, model, point_cloud_samples, config=None, shape_parallel=False, batch_size=32, generator=None,
                          device=None):
    """Shape latents of many condition shapes -- reference utils/shape.py:240-284 (centre every cloud on its mean, express the
    bounding box in that frame, stack batches of `batch_size`, encode).

    surfaces: a list whose items are   (verts [V,3], faces [F,3])           a triangle mesh (what get_mesh returns), or
                                        dict(coords=[n,3], atomic_numbers=[n]) a molecule (mesh-free surface), or
                                        dict(points=[P,3], bounds=[2,3])       an already sampled surface cloud.
    model:    the drop-in PointCloud_AE (its encoder runs in the CUDA library); clouds are encoded in chunks of `batch_size`
              (the DGCNN blocks' BatchNorm uses batch statistics per chunk, exactly like the reference's per-batch encode loop
              when shape_parallel is on; with shape_parallel off the reference encodes ALL clouds in one batch -- pass
              batch_size=len(surfaces) to get those statistics).
    Returns (zs [B,latent,3] on the CPU like the reference, bounds [B,3,2] (axis x (min | max), centred frame),
    [batches of clouds [b,1,P,3]], centers [B,3])."""
    if device is None:
        enc = getattr(model, 'encoder', None)
        device = next((p.device for p in enc.conv_pos.parameters()), torch.device('cuda')) if hasattr(enc, 'conv_pos') else torch.device('cpu')
    clouds, centers, bounds = [], [], []
    for s in surfaces:
        if isinstance(s, dict) and 'points' in s:
            pc = torch.as_tensor(s['points'], dtype=torch.float32, device=device)
            bd = torch.as_tensor(s['bounds'], dtype=torch.float32, device=device)
        elif isinstance(s, dict):
            pc, bd = sample_molecular_surface(torch.as_tensor(s['coords'], dtype=torch.float32, device=device), s['atomic_numbers'],
                                              point_cloud_samples, generator=generator)
        else:
            verts = torch.as_tensor(s[0], dtype=torch.float32, device=device)
            pc = sample_points_from_mesh(verts, s[1], point_cloud_samples, generator=generator)
            bd = mesh_bounds(verts)
        c = pc.mean(dim=0)
        clouds.append(pc - c)
        centers.append(c)
        bounds.append((bd - c).transpose(0, 1))          # [3, 2] = (bound.T - center).T of the reference (:262-264)
    batches = [torch.stack(clouds[i:i + batch_size]).unsqueeze(1) for i in range(0, len(clouds), batch_size)]
    zs = torch.cat([model.encoder(b).detach() for b in batches], dim=0).cpu() if batches else torch.zeros(0, 32, 3)
    return zs, torch.stack(bounds).cpu(), [b.cpu() for b in batches], torch.stack(centers).cpu()


def read_rdkit_pickle_coords(path):
    """Heavy-atom conformer coordinates of the molecules in a pickled list of RDKit binary mol blobs (the reference's
    data/MOSES2_test_mol.pkl) WITHOUT RDKit: a stub unpickler exposes each raw blob; the conformer is its trailing
    float32 [n,3] block, n = int32 at byte 20 (verified on all 1000 test molecules).  -> list of [n,3] float32 arrays."""
    import pickle

    class _Blob:
        def __init__(self, *a, **k):
            self.args = a

        def __setstate__(self, st):
            self.state = st

    class _U(pickle.Unpickler):
        def find_class(self, module, name):
            return _Blob

    with open(path, 'rb') as f:
        objs = _U(f).load()
    out = []
    for o in objs:
        blob = None
        for cand in list(getattr(o, 'args', ())) + [getattr(o, 'state', None)]:
            if isinstance(cand, (bytes, bytearray)):
                blob = bytes(cand)
        if blob is None:
            continue
        n = int(np.frombuffer(blob[20:24], dtype='<i4')[0])
        xyz = np.frombuffer(blob[-1 - 12 * n:-1], dtype='<f4').reshape(n, 3).copy()
        out.append(xyz)
    return out
