"""Host logic of the multi-GPU path (shapemol_b200/distributed.py) at world_size 2 on the gloo backend:
molecule sharding, global-atom Philox offsets and the single result gather."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from shapemol_b200 import distributed as D  # noqa: E402


def test_shard_ranges_cover_every_molecule_once():
    for n in (0, 1, 2, 7, 100, 5000):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and a <= b and c <= d
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= (n + world - 1) // world
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def test_shard_batch_atom_offsets_and_take_shard():
    g = torch.Generator().manual_seed(3)
    sizes = torch.randint(1, 28, (37,), generator=g)
    N = int(sizes.sum())
    pos, v, shape = torch.randn(N, 3, generator=g), torch.randint(0, 15, (N,), generator=g), torch.randn(37, 32, 3, generator=g)
    seen = 0
    for r in range(4):
        sh = D.take_shard(sizes, r, 4, pos, v, shape)
        assert sh['atom_offset'] == seen                      # Philox streams are keyed by the global atom index
        n = int(sh['sizes'].sum())
        assert sh['pos'].shape[0] == n == sh['batch'].numel()
        assert torch.equal(sh['pos'], pos[seen:seen + n]) and torch.equal(sh['v'], v[seen:seen + n])
        assert int(sh['batch'].max()) + 1 == sh['sizes'].numel() and torch.equal(torch.bincount(sh['batch']), sh['sizes'])
        assert torch.equal(sh['shape'], shape[sh['mol_offset']:sh['mol_offset'] + sh['sizes'].numel()])
        seen += n
    assert seen == N


def _worker(rank, world, port, sizes, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sizes_t = torch.tensor(sizes)
        N = int(sizes_t.sum())
        g = torch.Generator().manual_seed(11)
        pos, v = torch.randn(N, 3, generator=g), torch.randint(0, 15, (N,), generator=g)
        shape = torch.randn(len(sizes), 32, 3, generator=g)
        sh = D.take_shard(sizes_t, rank, world, pos, v, shape)
        # a "sampler" whose result depends only on the global atom index and the inputs: rank-count independent
        gidx = sh['atom_offset'] + torch.arange(sh['pos'].shape[0])
        out_pos = sh['pos'] * 2 + gidx[:, None].float()
        out_v = (sh['v'] + gidx) % 15
        full_pos, full_v = D.gather_results(out_pos, out_v, sizes_t)
        exp_pos = pos * 2 + torch.arange(N)[:, None].float()
        exp_v = (v + torch.arange(N)) % 15
        ok = torch.equal(full_pos, exp_pos) and torch.equal(full_v, exp_v) and full_v.dtype == torch.long
        try:   # a shard of the wrong size is rejected before any communication (same outcome on every rank)
            D.gather_results(torch.zeros(out_pos.shape[0] + 1, 3), torch.zeros(out_pos.shape[0] + 1), sizes_t)
            ok = False
        except ValueError:
            pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('sizes', [[5, 9, 27, 1, 13], [4], [3, 3, 3, 3]])
def test_gather_results_world2_gloo(sizes):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() + len(sizes) * 7) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
