"""Step time and live per-kernel-class launch times (CUDA events inside full evaluations) of the bf16 throughput mode.
    python tools/prof_step.py [--mols 16384] [--fixed-atoms 27] [--steps 10]
--fixed-atoms 0 = MOSES size prior (configs[1] shape).  Prints one line per kernel class and the step total."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mols', type=int, default=16384)
    ap.add_argument('--fixed-atoms', type=int, default=27)
    ap.add_argument('--k', type=int, default=32)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--precision', default='bf16')
    args = ap.parse_args()
    from shapemol_b200.engine import Sampler
    dev = torch.device('cuda', 0)
    model = bench.build_model(args.k, args.precision).to(dev).train()
    per = 64 if args.mols % 64 == 0 else 1
    sizes, batch, pos, v, shape = bench.make_workload(args.mols // per, per, 2021, args.fixed_atoms)
    B = int(sizes.numel())
    smp = Sampler(model._engine(), pos.to(dev), v.to(dev), batch.to(dev), shape.to(dev), num_steps=1000, noise='philox', seed=2021,
                  keep_traj=False, use_graph=True, n_mols=B)
    ms = bench.timed_graph_steps(smp, args.steps, 3, torch.cuda.synchronize)
    tiles, rows = bench.tiles_and_rows(sizes, args.k, split=False)
    tiles_s, _ = bench.tiles_and_rows(sizes, args.k)
    print('workload %d mols x %s  atoms %d  edges %d  tiles %d (X2H) / %d (gate, H2X)  rows used %.3f / %.3f' % (
        B, args.fixed_atoms or 'prior', int(sizes.sum()), rows, tiles, tiles_s, rows / (128.0 * tiles), rows / (128.0 * tiles_s)))
    print('step %.3f ms  %.0f mol-steps/s' % (ms, B / (ms * 1e-3)))
    if args.precision == 'bf16':
        kt = bench.prof_kernels(model._engine(), smp, ['edge_k', 'edge_v', 'edge_xv', 'node_pre', 'node_out', 'gate', 'knn', 'head'])
        tot = 0.0
        for c, (t, n) in kt.items():
            tot += t * n
            nt = {'edge_xv': tiles_s, 'gate': tiles_s, 'edge_k': 0.5 * (tiles + tiles_s)}.get(c, tiles)   # edge_k: mean of the two blocks
            print('  %-9s %8.4f ms x %2d = %7.3f ms (%4.1f%%)  %.1f ns/tile' % (c, t, n, t * n, 100 * t * n / ms, t * 1e6 / nt * 148))
        print('  other      %7.3f ms' % (ms - tot))


if __name__ == '__main__':
    main()
