"""The caller's side of the boundary: the reference's own `sample_diffusion_ligand` / `unbatch_v_traj`
(scripts/sample_diffusion.py:37-159), executed UNMODIFIED against the drop-in ScorePosNet3D.

scripts/sample_diffusion.py cannot be imported as a module here (rdkit, oddt, lmdb, torch_geometric ... are absent), so the
two function definitions are taken out of the script with `ast` at run time (from the reference checkout, or from the copy
oracle/stage_ref.py staged under oracle/_ref for the GPU box) and run in a namespace that supplies the handful of names they
use; `Batch.from_data_list` is a stand-in that honours the attributes the function reads.  What this pins: the 10-key return
dict of sample_diffusion is consumed by the reference's un-batching code exactly as it is (per-step lists, CPU / device
placement, dtypes), i.e. "the script runs unchanged" for the part of it that touches the model."""
import ast
import os
import time

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from test_gpu_parity import build_model

pytestmark = pytest.mark.gpu


def _script_source():
    for root in ('/root/reference', os.path.join(ROOT, 'oracle', '_ref')):
        p = os.path.join(root, 'scripts', 'sample_diffusion.py')
        if os.path.exists(p):
            return open(p).read()
    return None


class _Data:
    def __init__(self, n_atoms):
        self.ligand_element = torch.zeros(n_atoms, dtype=torch.long)

    def clone(self):
        return self


class _Batch:
    """torch_geometric.data.Batch stand-in: the attributes sample_diffusion_ligand reads (:60-61,:99,:110)."""

    def __init__(self, n_data, shape):
        self.ligand_smiles = ['C'] * n_data
        self.bound = torch.zeros(n_data * 3, 2)
        self.shape_emb = shape.repeat(n_data, 1, 1).reshape(-1, 3)

    @classmethod
    def from_data_list(cls, data_list, exclude_keys=None, follow_batch=None):
        return cls(len(data_list), cls.shape)

    def to(self, device):
        self.bound, self.shape_emb = self.bound.to(device), self.shape_emb.to(device)
        return self


def test_reference_sample_diffusion_ligand_runs_unchanged(cuda_lib):
    src = _script_source()
    if src is None:
        pytest.skip('neither /root/reference nor oracle/_ref holds scripts/sample_diffusion.py')
    tree = ast.parse(src)
    fns = {n.name: ast.get_source_segment(src, n) for n in tree.body if isinstance(n, ast.FunctionDef)}
    fx = load_golden('forward_k32_eval.pt')
    model = build_model(fx, 'bf16', training=True)
    import models.molopt_score_model as msm       # the drop-in (build_model installed it)
    from tqdm.auto import tqdm
    _Batch.shape = fx['shape'][:1]
    ns = {'np': np, 'torch': torch, 'tqdm': tqdm, 'time': time, 'Batch': _Batch, 'FOLLOW_BATCH': (), 'pdb': None,
          'log_sample_categorical': msm.log_sample_categorical, 'scatter_sum': None}
    exec(fns['unbatch_v_traj'], ns)
    exec(fns['sample_diffusion_ligand'], ns)
    sizes = [21, 9, 27, 17, 25]
    it = iter([sizes[:3], sizes[3:]])
    torch.manual_seed(0)
    num_steps = 6
    out = ns['sample_diffusion_ligand'](model, _Data(20), num_samples=5, batch_size=3, device='cuda:0', num_steps=num_steps,
                                        center_pos_mode='none', sample_func=lambda n: next(it), sample_num_atoms='size')
    pos, v, pos_traj, v_traj, v0_traj, vt_traj, time_list, pos_cond_traj, v_cond_traj = out
    assert len(pos) == 5 and [p.shape for p in pos] == [(n, 3) for n in sizes] and pos[0].dtype == np.float64
    assert [t.shape for t in pos_traj] == [(num_steps, n, 3) for n in sizes]
    assert [t.shape for t in v_traj] == [(num_steps, n) for n in sizes] and v_traj[0].dtype == np.int64
    assert [t.shape for t in v0_traj] == [(num_steps, n, 15) for n in sizes]
    assert [t.shape for t in vt_traj] == [(num_steps, n, 15) for n in sizes]
    assert [t.shape for t in pos_cond_traj] == [(num_steps, n, 3) for n in sizes]
    assert [t.shape for t in v_cond_traj] == [(num_steps, n, 15) for n in sizes]
    assert len(time_list) == 2 and all(np.isfinite(p).all() for p in pos)
    # the final state is the last trajectory entry, as evaluate_diffusion_sim.py assumes (eval_step = -1)
    assert all(np.array_equal(pt[-1], p) for pt, p in zip(pos_traj, pos))
    assert all(np.array_equal(vt[-1], vv) for vt, vv in zip(v_traj, v))
