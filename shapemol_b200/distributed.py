"""Multi-GPU sampling: molecules are independent units (kNN edges never cross molecules,
models/uni_transformer.py:468), so the path shards by molecule with NO per-step communication and ONE gather of
the final states.

One process per GPU (torchrun).  Rank r owns the contiguous block of molecules `shard_range(B, r, G)`; its Philox
noise streams are keyed by GLOBAL atom index (`atom_offset`), so with eval-mode BatchNorm the gathered result does
not depend on the number of ranks.  With train-mode BatchNorm (what scripts/sample_diffusion.py runs) a rank's shard
is its BatchNorm batch -- the reference run with batch_size = shard size.

The host logic here is backend-agnostic: tests/test_distributed_cpu.py runs it with gloo at world_size 2.
"""
import torch
import torch.distributed as dist


def shard_range(n_mols, rank, world):
    """Contiguous block [lo, hi) of molecule indices owned by `rank` (blocks of ceil(B / G))."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('bad rank %d / world %d' % (rank, world))
    per = (n_mols + world - 1) // world
    lo = min(n_mols, rank * per)
    return lo, min(n_mols, lo + per)


def shard_batch(sizes, rank, world):
    """sizes: atoms per molecule (1-D tensor / list).  Returns (mol_lo, mol_hi, atom_lo, atom_hi)."""
    sizes = torch.as_tensor(sizes, dtype=torch.long)
    lo, hi = shard_range(int(sizes.numel()), rank, world)
    ptr = torch.zeros(sizes.numel() + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(sizes, 0)
    return lo, hi, int(ptr[lo]), int(ptr[hi])


def take_shard(sizes, rank, world, pos, v, shape):
    """Slices a full problem (pos [N,3], v [N], shape [B,32,3]) down to this rank's molecules.
    Returns dict(pos, v, batch (local molecule index per atom), shape, sizes, atom_offset, mol_offset)."""
    sizes = torch.as_tensor(sizes, dtype=torch.long)
    lo, hi, a_lo, a_hi = shard_batch(sizes, rank, world)
    local = sizes[lo:hi]
    batch = torch.repeat_interleave(torch.arange(hi - lo), local)
    return dict(pos=pos[a_lo:a_hi], v=v[a_lo:a_hi], batch=batch, shape=shape[lo:hi], sizes=local, atom_offset=a_lo, mol_offset=lo)


def gather_results(pos, v, sizes, group=None):
    """The path's only collective: every rank contributes its final (pos [n_r,3] f32, v [n_r] int) and receives the
    full (pos [N,3], v [N]) in global molecule order.  `sizes` is the GLOBAL atoms-per-molecule vector.
    One all_gather of a fixed-size packed int32 buffer [max_shard_atoms, 4] = (x, y, z bit patterns | v)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = torch.as_tensor(sizes, dtype=torch.long)
    spans = [shard_batch(sizes, r, world)[2:] for r in range(world)]
    n_r = spans[rank][1] - spans[rank][0]
    if pos.shape[0] != n_r or v.shape[0] != n_r:
        raise ValueError('rank %d holds %d atoms, its shard has %d' % (rank, pos.shape[0], n_r))
    cap = max(1, max(hi - lo for lo, hi in spans))
    buf = torch.zeros(cap, 4, dtype=torch.int32, device=pos.device)
    if n_r:
        buf[:n_r, :3] = pos.detach().to(torch.float32).contiguous().view(torch.int32)
        buf[:n_r, 3] = v.to(torch.int32)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    pos_all = torch.cat([o[:hi - lo, :3] for o, (lo, hi) in zip(out, spans)]).contiguous().view(torch.float32)
    v_all = torch.cat([o[:hi - lo, 3] for o, (lo, hi) in zip(out, spans)]).to(torch.long)
    return pos_all, v_all


def sample_sharded(model, pos, v, sizes, shape, num_steps=None, seed=2021, group=None, noise='philox'):
    """Full problem on every rank's host -> each rank samples its own molecules on its own GPU -> gathered
    (pos [N,3], v [N]) on every rank.  `model` is the drop-in ScorePosNet3D already on this rank's device."""
    from .engine import Sampler
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = next(model.parameters()).device
    sh = take_shard(sizes, rank, world, pos, v, shape)
    sampler = Sampler(model._engine(), sh['pos'].to(dev), sh['v'].to(dev), sh['batch'].to(dev), sh['shape'].to(dev),
                      num_steps=num_steps, noise=noise, seed=seed, atom_offset=sh['atom_offset'], keep_traj=False,
                      n_mols=int(sh['sizes'].numel()))
    p, vv = sampler.run()
    return gather_results(p, vv, sizes, group=group)
