"""Multi-GPU correctness of the molecule-sharded sampling path, run under torchrun on N >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/check_sharded.py
Every rank samples its shard (shapemol_b200.distributed.sample_sharded); rank 0 also samples the whole problem on one GPU.
With eval-mode BatchNorm and Philox noise keyed by the global atom index the gathered result must be IDENTICAL to the
single-GPU result, whatever the number of ranks."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from conftest import load_golden
    from test_gpu_parity import build_model
    from shapemol_b200 import distributed as D
    from shapemol_b200.engine import Sampler
    fx = load_golden('forward_k32_eval.pt')
    ok = True
    for precision in ('bf16', 'bf16x3'):
        m = build_model(fx, precision, training=False).to('cuda')
        g = torch.Generator().manual_seed(5)
        B = 203                                            # not divisible by the rank count
        sizes = torch.randint(9, 28, (B,), generator=g)
        N = int(sizes.sum())
        pos, v = torch.randn(N, 3, generator=g), torch.randint(0, 15, (N,), generator=g)
        shape = 0.07 * torch.randn(B, 32, 3, generator=g)
        p_all, v_all = D.sample_sharded(m, pos, v, sizes, shape, num_steps=8, seed=77)
        if rank == 0:
            batch = torch.repeat_interleave(torch.arange(B), sizes).cuda()
            s = Sampler(m._engine(), pos.cuda(), v.cuda(), batch, shape.cuda(), num_steps=8, noise='philox', seed=77, atom_offset=0,
                        keep_traj=False, n_mols=B)
            p1, v1 = s.run()
            same = bool(torch.equal(p_all, p1)) and bool(torch.equal(v_all, v1.long()))
            print('%s: %d ranks, %d molecules, %d atoms: sharded == single GPU: %s  (max |dpos| %.3g)'
                  % (precision, world, B, N, same, float((p_all - p1).abs().max())), flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == '__main__':
    main()
