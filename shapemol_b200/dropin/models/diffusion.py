"""Variance schedules (host-side constants, float64 numpy) -- same functions and semantics as the
reference's models/diffusion.py:4-48."""
import numpy as np


def cosine_beta_schedule(timesteps, s=0.008):
    n = timesteps + 1
    grid = np.linspace(0, n, n)
    acp = np.cos(((grid / n) + s) / (1 + s) * np.pi * 0.5) ** 2
    acp = acp / acp[0]
    return np.clip(1 - acp[1:] / acp[:-1], 0, 0.999)


def get_beta_schedule(beta_schedule, num_diffusion_timesteps, **kwargs):
    kw = {k: float(v) for k, v in kwargs.items()}
    T = num_diffusion_timesteps
    if beta_schedule == 'quad':
        betas = np.linspace(kw['beta_start'] ** 0.5, kw['beta_end'] ** 0.5, T, dtype=np.float64) ** 2
    elif beta_schedule == 'linear':
        betas = np.linspace(kw['beta_start'], kw['beta_end'], T, dtype=np.float64)
    elif beta_schedule == 'sigmoid':
        s = kw.get('s', 3)
        ramp = np.linspace(-s, s, T)
        betas = (1 / (np.exp(-ramp) + 1)) * (kw['beta_end'] - kw['beta_start']) + kw['beta_start']
    elif beta_schedule == 'cosine':
        betas = cosine_beta_schedule(T, s=kw.get('s', 0.008))
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (T,)
    return betas
