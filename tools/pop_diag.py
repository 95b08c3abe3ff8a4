"""Diagnostic: final-state statistics of the 1000-step reverse process, CUDA path (several modes) vs tests/golden/population.pt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
from oracle import shapemol_oracle as orc  # noqa: E402
from test_gpu_distribution import _molecule_stats  # noqa: E402
from test_host_cpu import make_dropin  # noqa: E402
from test_gpu_parity import batch_of  # noqa: E402
from conftest import manifest_shapes  # noqa: E402
import synth  # noqa: E402

fx = torch.load(os.path.join(ROOT, 'tests', 'golden', 'population.pt'))
sizes, shape, steps = fx['sizes'], fx['shape'], fx['steps']
B = len(sizes)
mol_ptr = orc.mol_ptr_from_sizes(sizes)
N = int(mol_ptr[-1])
g = torch.Generator().manual_seed(5)
ref = (1.5 * torch.randn(B * 20, 3, generator=g)).double()
ref_ptr = torch.arange(0, B * 20 + 1, 20)


def show(name, pos, v):
    s = _molecule_stats(pos, v, mol_ptr, ref, ref_ptr)
    print('%-28s rg %.3f +- %.3f  pair %.3f  nn %.3f  rocs %.4f  types %s' % (name, float(s['rg'].mean()), float(s['rg'].std()), float(s['pair'].mean()),
                                                                             float(s['nn'].mean()), float(s['rocs'].mean()), [int(x) for x in s['types']]), flush=True)


for i, r in enumerate(fx['runs']):
    show('reference run %d' % i, r['pos'], r['v'])
sd = synth.synth_state_dict(manifest_shapes(), fx['seed'])
import statistics
for precision, noise in (('bf16x3', 'philox'), ('bf16x3', 'torch'), ('bf16', 'philox')):
    vals = []
    for seed in range(8):
        m, _ = make_dropin(knn=fx['k'], num_diffusion_timesteps=steps)
        m.load_state_dict(sd, strict=False)
        m = m.cuda().train()
        m.smb_precision, m.smb_noise, m.smb_seed, m.smb_keep_traj = precision, noise, 100 + seed, False
        gg = torch.Generator().manual_seed(700 + seed)
        pos1, v1 = torch.randn(N, 3, generator=gg), torch.randint(0, 15, (N,), generator=gg)
        torch.manual_seed(900 + seed)
        r = m.sample_diffusion(pos1.cuda(), v1.cuda(), batch_of(sizes), shape.view(-1, 3).cuda(), num_steps=steps, center_pos_mode='none')
        st = _molecule_stats(r['pos'], r['v'], mol_ptr, ref, ref_ptr)
        vals.append((float(st['rg'].mean()), float(st['pair'].mean()), float(st['nn'].mean())))
    print('%s %s: rg means %s | mean %.3f sd %.3f' % (precision, noise, ' '.join('%.3f' % v[0] for v in vals), statistics.mean(v[0] for v in vals),
                                                     statistics.stdev(v[0] for v in vals)), flush=True)
