"""Vector-neuron parameter holders (reference models/shape_vn_layers.py:41-110)."""
import torch.nn as nn

EPS = 1e-6


class VNBatchNorm(nn.Module):
    def __init__(self, num_features, dim):
        super().__init__()
        self.dim = dim
        self.bn = nn.BatchNorm1d(num_features) if dim in (3, 4) else nn.BatchNorm2d(num_features)


class VNLinearLeakyReLU(nn.Module):
    def __init__(self, in_channels, out_channels, dim=5, share_nonlinearity=False, negative_slope=0.2,
                 use_batchnorm=True):
        super().__init__()
        self.dim = dim
        self.negative_slope = negative_slope
        self.map_to_feat = nn.Linear(in_channels, out_channels, bias=False)
        self.use_batchnorm = use_batchnorm
        if use_batchnorm:
            self.batchnorm = VNBatchNorm(out_channels, dim=dim)
        self.map_to_dir = nn.Linear(in_channels, 1 if share_nonlinearity else out_channels, bias=False)

    def forward(self, x):
        raise RuntimeError('VNLinearLeakyReLU is evaluated inside the shapemol_b200 CUDA kernels')
