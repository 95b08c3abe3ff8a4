"""Deterministic synthetic weights shared by the golden generator and the tests.

The trained diffusion checkpoint is absent from the reference checkout, so parity is
pinned with synthetic weights: every tensor of a state_dict is filled from a CPU torch.Generator
seeded with (seed, crc32(key)) -- the fixtures therefore store only inputs and outputs."""
import zlib

import torch


def _fill(key, shape, seed):
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    if key.endswith('num_batches_tracked'):
        return torch.zeros(shape, dtype=torch.long)
    if key.endswith('running_mean'):
        return 0.6 + 0.1 * torch.randn(shape, generator=g)
    if key.endswith('running_var'):
        return 0.05 + 0.1 * torch.rand(shape, generator=g)
    is_norm = ('.net.1.' in key) or ('.bn.' in key)
    if is_norm and key.endswith('weight'):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if is_norm and key.endswith('bias'):
        return 0.1 * torch.randn(shape, generator=g)
    if key.endswith('bias'):
        return 0.1 * (2 * torch.rand(shape, generator=g) - 1)
    fan_in = shape[-1] if len(shape) > 1 else 1
    bound = 1.0 / (fan_in ** 0.5)
    return bound * (2 * torch.rand(shape, generator=g) - 1)


def synth_state_dict(shapes, seed, skip_non_synth=True):
    """shapes: dict key -> tuple.  Schedule tables and RBF offsets (is_non_synth) are left out."""
    out = {}
    for key, shape in shapes.items():
        if skip_non_synth and is_non_synth(key):
            continue
        out[key] = _fill(key, tuple(shape), seed)
    return out


SCHEDULE_KEYS = ('loss_pos_step_weight', 'betas', 'alphas_cumprod', 'alphas_cumprod_prev', 'sqrt_alphas_cumprod',
                 'sqrt_one_minus_alphas_cumprod', 'sqrt_recip_alphas_cumprod', 'sqrt_recipm1_alphas_cumprod',
                 'posterior_mean_c0_coef', 'posterior_mean_ct_coef', 'posterior_var', 'posterior_logvar',
                 'log_alphas_v', 'log_one_minus_alphas_v', 'log_alphas_cumprod_v', 'log_one_minus_alphas_cumprod_v')
NON_SYNTH = SCHEDULE_KEYS + ('distance_expansion.offset',)


def is_non_synth(key):
    return key in SCHEDULE_KEYS or key.endswith('distance_expansion.offset')


def molecule_like_positions(sizes, seed, spread=1.6):
    """Random compact 3-D point sets with molecule-like spacing (no two atoms closer than ~0.9 A)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in sizes:
        pts = []
        while len(pts) < n:
            c = torch.randn(3, generator=g) * spread * (n ** (1 / 3)) * 0.6
            if all(torch.norm(c - p) > 0.9 for p in pts):
                pts.append(c)
        out.append(torch.stack(pts))
    return torch.cat(out, 0).float()
