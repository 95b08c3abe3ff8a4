"""Parameter holders with the reference's names (models/common.py).  The arithmetic runs in the CUDA
library; these modules only own the tensors so that state_dict()/load_state_dict() are identical."""
import math

import torch
import torch.nn as nn

RBF_CENTRES = [0, 1, 1.25, 1.5, 1.75, 2, 2.25, 2.5, 2.75, 3, 3.5, 4, 4.5, 5, 5.5, 6, 7, 8, 9, 10]


def _kernel_only(name):
    raise RuntimeError('%s is evaluated inside the shapemol_b200 CUDA kernels; it has no eager path' % name)


class GaussianSmearing(nn.Module):
    """20 fixed centres, coeff -0.5 (reference models/common.py:11-28)."""

    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        self.start, self.stop, self.num_gaussians = start, stop, num_gaussians
        offset = torch.tensor(RBF_CENTRES, dtype=torch.float32)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer('offset', offset)

    def forward(self, dist):
        _kernel_only('GaussianSmearing')


class ShiftedSoftplus(nn.Module):
    def __init__(self):
        super().__init__()
        self.shift = math.log(2.0)

    def forward(self, x):
        _kernel_only('ShiftedSoftplus')


class MLP(nn.Module):
    """Linear -> LayerNorm -> ReLU -> Linear, stored as `net.{0,1,3}` like the reference (:47-67)."""

    def __init__(self, in_dim, out_dim, hidden_dim, num_layer=2, norm=True, act_fn='relu', act_last=False):
        super().__init__()
        if num_layer != 2 or not norm or act_fn != 'relu' or act_last:
            raise NotImplementedError('only the 2-layer LayerNorm/ReLU MLP of the shipped config is built')
        self.net = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(),
                                 nn.Linear(hidden_dim, out_dim))

    def forward(self, x):
        _kernel_only('MLP')
