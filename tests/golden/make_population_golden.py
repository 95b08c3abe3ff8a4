"""Final state of a FULL 1000-step reverse process of the UNMODIFIED reference (models/molopt_score_model.py
sample_diffusion, train-mode BatchNorm like the trajectory fixture, synthetic weights -- the trained checkpoint is not in the
reference checkout), for the distribution-level parity test (tests/test_gpu_distribution.py).

Eight independent runs (different initial noise and torch seeds): train-mode BatchNorm couples the 48 molecules of a run, so the
run is the independent unit of the comparison and the run-to-run spread of the reference calibrates it.

Run in the build container only (about an hour on 8 cores; resumes from an existing population.pt):   python tests/golden/make_population_golden.py
Output: tests/golden/population.pt
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from make_golden import build_model, batch_of  # noqa: E402

torch.set_num_threads(8)


def main():
    msm, spm, ED = ref_loader.load()
    seed, k, steps, B = 11, 32, 1000, 48
    m, _ = build_model(msm, ED, seed, True, knn=k)
    g = torch.Generator().manual_seed(2024)
    sizes = torch.randint(9, 28, (B,), generator=g).tolist()
    N = sum(sizes)
    shape = 0.07 * torch.randn(B, 32, 3, generator=g)
    path = os.path.join(HERE, 'population.pt')
    out = torch.load(path) if os.path.exists(path) else dict(seed=seed, k=k, steps=steps, sizes=sizes, shape=shape, runs=[])
    assert out['sizes'] == sizes and torch.equal(out['shape'], shape)
    done = {r['noise_seed'] for r in out['runs']}
    for run, noise_seed in enumerate(range(5, 13)):
        if noise_seed in done:
            continue
        gg = torch.Generator().manual_seed(100 + noise_seed)
        pos0 = torch.randn(N, 3, generator=gg)
        v0 = torch.randint(0, 15, (N,), generator=gg)
        torch.manual_seed(noise_seed)
        t0 = time.time()
        with torch.no_grad():
            r = m.sample_diffusion(init_ligand_pos=pos0, init_ligand_v=v0, batch_ligand=batch_of(sizes),
                                   ligand_shape=shape.view(-1, 3), num_steps=steps, center_pos_mode='none')
        print('run %d: %.0f s, finite %s' % (run, time.time() - t0, bool(torch.isfinite(r['pos']).all())), flush=True)
        out['runs'].append(dict(noise_seed=noise_seed, pos0=pos0, v0=v0, pos=r['pos'].clone(), v=r['v'].clone()))
        torch.save(out, path)


if __name__ == '__main__':
    main()
