"""Drop-in PointCloud_AE / VN_DGCNN_Encoder (reference models/shape_pointcloud_modelAE.py:105-255).

Same class names, constructor, custom `.to(device)` and state_dict keys as the reference, so
`utils/shape.py:226-238` (build_point_shapeAE_model: construct, `.to`, `load_state_dict(strict=True)`)
and `model.encoder(x[B,1,P,3]) -> [B,latent,3]` (`utils/shape.py:280`) run unchanged.  The encoder's
arithmetic runs in libshapemol_b200 (smb_vn_dgcnn_encode); there is no eager / CPU fallback.

Reference quirks kept on purpose:
  * `VN_DGCNN_Encoder.blocks` is a plain Python list: the four DGCNN blocks are NOT in the state_dict,
    keep their constructor initialisation, ignore `.train()/.eval()` and always use batch statistics;
  * `conv_pos` / `conv_c` follow `module.training`.
The shape decoder (`generator`) is outside the hot path: its parameters are held for
`load_state_dict(strict=True)`; calling it raises.
"""
import ctypes as C

import torch
import torch.nn as nn

from models.shape_vn_layers import VNLinearLeakyReLU


class VNLinear(nn.Module):
    """parameter holder of models/shape_vn_layers.py:8-17"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.map_to_feat = nn.Linear(in_channels, out_channels, bias=False)


class DecoderInner(nn.Module):
    """Parameter holder of the occupancy / signed-distance decoder (reference :21-103); out of scope."""

    def __init__(self, dim=3, z_dim=128, hidden_size=128, layer_num=4, loss_type='occupancy'):
        super().__init__()
        self.z_dim = z_dim
        self.layer_num = layer_num
        if z_dim > 0:
            self.z_in = VNLinear(z_dim, z_dim)
        self.fc_in = nn.Linear(z_dim * 2 + 1, hidden_size)
        self.fc_out = nn.Linear(hidden_size, 1)
        self.loss_type = loss_type

    def to(self, device):
        if self.z_dim > 0:
            self.z_in = self.z_in.to(device)
        self.fc_in = self.fc_in.to(device)
        self.fc_out = self.fc_out.to(device)
        return self

    def forward(self, *a, **k):
        raise RuntimeError('the shape decoder is outside the shapemol_b200 hot path (SURVEY 2); use the reference module')


class VN_DGCNN_Encoder(nn.Module):

    def __init__(self, hidden_dim, latent_dim, layer_num, num_k):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.latent_dim = latent_dim
        self.layer_num = layer_num
        self.num_k = num_k
        self.blocks = []          # plain list, exactly like the reference (:212-218)
        self.conv_pos = VNLinearLeakyReLU(2, hidden_dim)
        final_input_dim = 0
        for _ in range(layer_num):
            self.blocks.append(VNLinearLeakyReLU(2 * hidden_dim, hidden_dim))
            final_input_dim += hidden_dim
        self.conv_c = VNLinearLeakyReLU(final_input_dim, latent_dim, dim=4, share_nonlinearity=True)
        self._ws = None

    def to(self, device):
        self.conv_pos = self.conv_pos.to(device)
        for i in range(len(self.blocks)):
            self.blocks[i] = self.blocks[i].to(device)
        self.conv_c = self.conv_c.to(device)
        return self

    def _weights(self, dev):
        from shapemol_b200 import _lib
        w = _lib.EncoderWeights()
        w.hidden, w.latent, w.n_blocks, w.num_k = self.hidden_dim, self.latent_dim, self.layer_num, self.num_k

        def dp(t):
            if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                raise _lib.SmbError('encoder tensors must be contiguous fp32 on %s (call .to(device) first)' % dev)
            return t.data_ptr()
        for name, layer in (('conv_pos', self.conv_pos), ('conv_c', self.conv_c)):
            bn = layer.batchnorm.bn
            setattr(w, name + '_feat', dp(layer.map_to_feat.weight))
            setattr(w, name + '_dir', dp(layer.map_to_dir.weight))
            setattr(w, name + '_bn_w', dp(bn.weight))
            setattr(w, name + '_bn_b', dp(bn.bias))
            setattr(w, name + '_bn_rm', dp(bn.running_mean))
            setattr(w, name + '_bn_rv', dp(bn.running_var))
        for i, blk in enumerate(self.blocks):
            bn = blk.batchnorm.bn
            w.block_feat[i], w.block_dir[i] = dp(blk.map_to_feat.weight), dp(blk.map_to_dir.weight)
            w.block_bn_w[i], w.block_bn_b[i] = dp(bn.weight), dp(bn.bias)
            w.block_bn_rm[i], w.block_bn_rv[i] = dp(bn.running_mean), dp(bn.running_var)
        # conv_pos / conv_c follow their own training flag (the reference never toggles them apart)
        if self.conv_pos.training != self.conv_c.training:
            raise _lib.SmbError('conv_pos and conv_c must be in the same train/eval mode')
        w.training = int(self.conv_pos.training)
        return w

    def forward(self, input):
        """input [B, 1, P, 3] -> latent [B, latent_dim, 3]   (reference :231-255)"""
        from shapemol_b200 import _lib
        lib = _lib.load()
        if input.dim() != 4 or input.size(1) != 1 or input.size(3) != 3:
            raise ValueError('expected input of shape [B, 1, P, 3], got %s' % (tuple(input.shape),))
        dev = input.device
        if dev.type != 'cuda':
            raise _lib.SmbError('shapemol_b200 runs on CUDA devices only (got %s); there is no CPU path' % dev)
        B, P = int(input.size(0)), int(input.size(2))
        clouds = input.detach().to(torch.float32).reshape(B, P, 3).contiguous()
        w = self._weights(dev)
        need = lib.smb_encoder_workspace_bytes(C.byref(w), B, P)
        if need == 0:
            _lib.check(-1, 'smb_encoder_workspace_bytes')
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        latent = torch.empty(B, self.latent_dim, 3, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.no_grad():
            _lib.check(lib.smb_vn_dgcnn_encode(C.byref(w), clouds.data_ptr(), B, P, latent.data_ptr(), self._ws.data_ptr(),
                                               self._ws.numel(), stream), 'smb_vn_dgcnn_encode')
            if w.training:
                for layer in [self.conv_pos, self.conv_c] + list(self.blocks):
                    layer.batchnorm.bn.num_batches_tracked += 1
            else:
                for layer in self.blocks:          # the blocks are always in train mode
                    layer.batchnorm.bn.num_batches_tracked += 1
        return latent


class PointCloud_AE(nn.Module):

    def __init__(self, config):
        super().__init__()
        if config.encoder != 'VN_DGCNN':
            raise ValueError('only the VN_DGCNN encoder is built (got %r)' % (config.encoder,))
        self.encoder = VN_DGCNN_Encoder(config.hidden_dim, config.latent_dim, config.layer_num, config.num_k)
        self.generator = DecoderInner(config.point_dim, config.latent_dim, config.hidden_dim, config.layer_num, config.loss_type)
        self.loss_type = config.loss_type

    def to(self, device):
        self.encoder = self.encoder.to(device)
        self.generator = self.generator.to(device)
        return self

    def forward(self, inputs, z_vector, point_coord, is_training=False):
        """reference :121-135: encode when inputs are given; decoding is out of scope."""
        if is_training or inputs is not None:
            z_vector = self.encoder(inputs)
        if is_training or (z_vector is not None and point_coord is not None):
            raise RuntimeError('the shape decoder is outside the shapemol_b200 hot path (SURVEY 2); use the reference module')
        return z_vector, None
