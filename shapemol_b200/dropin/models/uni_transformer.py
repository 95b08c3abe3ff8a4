"""Parameter tree of the equivariant network, names identical to the reference
(models/uni_transformer.py).  Evaluation happens in libshapemol_b200 (see shapemol_b200/engine.py)."""
import torch.nn as nn

from models.common import GaussianSmearing, MLP
from models.shape_vn_layers import VNLinearLeakyReLU


class BaseX2HAttLayer(nn.Module):
    """reference :16-46"""

    def __init__(self, input_dim, hidden_dim, output_dim, shape_dim, n_heads, edge_feat_dim, r_feat_dim,
                 act_fn='relu', norm=True, shape_mode='attention', topo_emb_type='topo_layer', out_fc=True):
        super().__init__()
        kv = input_dim * 2 + edge_feat_dim + r_feat_dim + shape_dim
        self.n_heads = n_heads
        self.hk_func = MLP(kv, output_dim, hidden_dim, norm=norm, act_fn=act_fn)
        self.hv_func = MLP(kv, output_dim, hidden_dim, norm=norm, act_fn=act_fn)
        self.hq_func = MLP(input_dim, output_dim, hidden_dim, norm=norm, act_fn=act_fn)
        self.node_output = MLP(2 * hidden_dim, hidden_dim, hidden_dim, norm=norm, act_fn=act_fn)


class BaseH2XAttLayer(nn.Module):
    """reference :93-119"""

    def __init__(self, input_dim, hidden_dim, output_dim, shape_dim, n_heads, edge_feat_dim, r_feat_dim,
                 act_fn='relu', norm=True, shape_mode='attention_residue', topo_emb_type='topo_layer'):
        super().__init__()
        kv = input_dim * 2 + edge_feat_dim + r_feat_dim + shape_dim
        self.n_heads = n_heads
        self.xk_func = MLP(kv, output_dim, hidden_dim, norm=norm, act_fn=act_fn)
        self.xv_func = MLP(kv, n_heads, hidden_dim, norm=norm, act_fn=act_fn)
        self.xq_func = MLP(input_dim, output_dim, hidden_dim, norm=norm, act_fn=act_fn)
        self.shape_linear = VNLinearLeakyReLU(n_heads + shape_dim + 1, n_heads, dim=4)


class EquivariantShapeEmbLayer(nn.Module):
    """Constructed but never called by the reference (:165-174, :393); kept for state_dict parity."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.hidden_layer = VNLinearLeakyReLU(input_dim, output_dim, dim=4)


class InvariantShapeEmbLayer(nn.Module):
    """reference :176-189"""

    def __init__(self, input_dim, output_dim, act_fn='relu', norm=True):
        super().__init__()
        self.hidden_layer = MLP(input_dim, output_dim, input_dim, norm=norm, act_fn=act_fn)


class AttentionLayerO2TwoUpdateNodeGeneral(nn.Module):
    """reference :242-287"""

    def __init__(self, hidden_dim, n_heads, num_r_gaussian, edge_feat_dim, shape_dim, act_fn='relu', norm=True,
                 num_x2h=1, num_h2x=1, r_max=10., topo_emb_type=None, x2h_out_fc=True):
        super().__init__()
        self.distance_expansion = GaussianSmearing(0., r_max, num_gaussians=num_r_gaussian)
        self.x2h_layers = nn.ModuleList([
            BaseX2HAttLayer(hidden_dim, hidden_dim, hidden_dim, shape_dim, n_heads, edge_feat_dim,
                            r_feat_dim=num_r_gaussian, act_fn=act_fn, norm=norm, out_fc=x2h_out_fc)
            for _ in range(num_x2h)])
        self.h2x_layers = nn.ModuleList([
            BaseH2XAttLayer(hidden_dim, hidden_dim, hidden_dim, shape_dim, n_heads, edge_feat_dim,
                            r_feat_dim=num_r_gaussian, act_fn=act_fn, norm=norm)
            for _ in range(num_h2x)])


class UniTransformerO2TwoUpdateGeneral(nn.Module):
    """reference :336-393.  Configurations outside the fast path raise at construction (no fallback)."""

    def __init__(self, num_blocks, num_layers, hidden_dim, shape_dim, shape_latent_dim, n_heads=1, k=32,
                 num_r_gaussian=50, edge_feat_dim=0, num_node_types=8, act_fn='relu', norm=True,
                 cutoff_mode='radius', shape_coeff=0.25, ew_net_type='global', topo_emb_type='topo_layer',
                 r_feat_mode='basic', num_topo=8, num_init_x2h=1, num_init_h2x=0, num_x2h=1, num_h2x=1, r_max=10.,
                 x2h_out_fc=True, atom_enc_mode='add_aromatic', shape_type='pointAE_shape', sync_twoup=False):
        super().__init__()
        unsupported = []
        if num_blocks != 1: unsupported.append('num_blocks != 1')
        if topo_emb_type in ('topo_layer', 'topo_attr'): unsupported.append('topo_emb_type=%s' % topo_emb_type)
        if cutoff_mode != 'knn': unsupported.append('cutoff_mode=%s' % cutoff_mode)
        if ew_net_type != 'global': unsupported.append('ew_net_type=%s' % ew_net_type)
        if edge_feat_dim != 0: unsupported.append('edge_feat_dim != 0')
        if num_r_gaussian != 20: unsupported.append('num_r_gaussian != 20')
        if num_x2h != 1 or num_h2x != 1: unsupported.append('num_x2h/num_h2x != 1')
        if sync_twoup: unsupported.append('sync_twoup')
        if shape_type != 'pointAE_shape': unsupported.append('shape_type=%s' % shape_type)
        if act_fn != 'relu' or not norm: unsupported.append('act_fn/norm')
        if shape_dim != 32 or shape_latent_dim != 32: unsupported.append('shape_dim != 32')
        if unsupported:
            raise NotImplementedError('shapemol_b200 builds only the shipped hot path; unsupported: ' + ', '.join(unsupported))
        self.num_blocks, self.num_layers, self.hidden_dim, self.n_heads, self.k = num_blocks, num_layers, hidden_dim, n_heads, k
        self.num_r_gaussian, self.edge_feat_dim, self.shape_dim = num_r_gaussian, edge_feat_dim, shape_dim
        self.cutoff_mode, self.ew_net_type, self.topo_emb_type = cutoff_mode, ew_net_type, topo_emb_type
        self.distance_expansion = GaussianSmearing(0., r_max, num_gaussians=num_r_gaussian)
        self.edge_pred_layer = MLP(num_r_gaussian, 1, hidden_dim)
        self.base_block = nn.ModuleList([
            AttentionLayerO2TwoUpdateNodeGeneral(hidden_dim, n_heads, num_r_gaussian, edge_feat_dim, shape_dim,
                                                 act_fn=act_fn, norm=norm, num_x2h=num_x2h, num_h2x=num_h2x,
                                                 r_max=r_max, topo_emb_type=topo_emb_type, x2h_out_fc=x2h_out_fc)
            for _ in range(num_layers)])
        self.invariant_shape_layer = InvariantShapeEmbLayer(shape_dim, shape_latent_dim)
        self.equivariant_shape_layer = EquivariantShapeEmbLayer(shape_dim, shape_latent_dim // 3)

    def forward(self, *args, **kwargs):
        raise RuntimeError('UniTransformerO2TwoUpdateGeneral is evaluated by ScorePosNet3D.forward through the CUDA library')
