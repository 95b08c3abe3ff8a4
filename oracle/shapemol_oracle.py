"""ORACLE -- test infrastructure, NOT product code.

A CPU restatement (torch CPU ops / numpy, no nn.Module tree of its own) of the reference's
reverse-diffusion denoising step and VN-DGCNN shape encoder.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this file; the product path
(shapemol_b200/) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  tests/golden/make_golden.py runs the UNMODIFIED reference modules
(/root/reference/models/*.py, imported through the shims in tests/golden/_shims) in the build
container and commits their outputs under tests/golden/*.pt; tests/test_oracle_golden.py checks
every function below against those fixtures.  The two third-party ops the reference calls but does
not vendor (torch_cluster.knn via torch_geometric.nn.knn_graph, torch_scatter.scatter_softmax /
scatter_sum; pinned only by ReadMe.md:15-17 to torch-cluster 1.6.0 / torch-scatter 2.0.9) are
restated from their published semantics -- for those two the reference holds no test, so their
parity is anchored on the reference's call sites (models/uni_transformer.py:468, :77, :80, :147,
:151) and stated as such in DESIGN.md.

Every function cites the reference file:line it follows.  Weights come in as a plain dict with the
reference's own state_dict key names.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-6
RBF_CENTRES = [0, 1, 1.25, 1.5, 1.75, 2, 2.25, 2.5, 2.75, 3, 3.5, 4, 4.5, 5, 5.5, 6, 7, 8, 9, 10]


# --------------------------------------------------------------------------------------
# schedules  (models/diffusion.py:4-48, models/molopt_score_model.py:188-234)
# --------------------------------------------------------------------------------------
def beta_schedule(kind, T, **kw):
    """models/diffusion.py:4-35 (float64 numpy)."""
    kw = {k: float(v) for k, v in kw.items()}
    if kind == 'sigmoid':
        s = kw.get('s', 3)
        b = np.linspace(-s, s, T)
        b = 1.0 / (np.exp(-b) + 1.0)
        return b * (kw['beta_end'] - kw['beta_start']) + kw['beta_start']
    if kind == 'cosine':
        # models/diffusion.py:38-48
        s = kw.get('s', 0.008)
        steps = T + 1
        x = np.linspace(0, steps, steps)
        ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
        ac = ac / ac[0]
        return np.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    if kind == 'linear':
        return np.linspace(kw['beta_start'], kw['beta_end'], T, dtype=np.float64)
    if kind == 'quad':
        return np.linspace(kw['beta_start'] ** 0.5, kw['beta_end'] ** 0.5, T, dtype=np.float64) ** 2
    raise NotImplementedError(kind)


def schedule_tables(T, schedule_pos, schedule_v):
    """The seven fp32 [T] tables the sampling loop reads (models/molopt_score_model.py:188-234)."""
    sp = dict(schedule_pos)
    betas = beta_schedule(sp.pop('beta_schedule'), T, **sp)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    c0 = betas * np.sqrt(ac_prev) / (1.0 - ac)
    ct = (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)
    # :218-220 -- posterior_var is first rounded to fp32 (to_torch_const), then log'ed in fp32
    pv32 = torch.from_numpy(post_var).float()
    logvar = np.log(np.append(pv32[1], pv32[1:]))
    sv = dict(schedule_v)
    betas_v = beta_schedule(sv.pop('beta_schedule'), T, **sv)
    la = np.log(1.0 - betas_v)
    lac = np.cumsum(la)

    def l1m(a):  # models/molopt_score_model.py:107-108
        return np.log(1 - np.exp(a) + 1e-40)

    def f(a):
        return torch.from_numpy(np.asarray(a)).float()
    return {
        'posterior_mean_c0_coef': f(c0), 'posterior_mean_ct_coef': f(ct), 'posterior_logvar': f(logvar),
        'log_alphas_v': f(la), 'log_one_minus_alphas_v': f(l1m(la)),
        'log_alphas_cumprod_v': f(lac), 'log_one_minus_alphas_cumprod_v': f(l1m(lac)),
    }


# --------------------------------------------------------------------------------------
# primitives  (models/common.py)
# --------------------------------------------------------------------------------------
def rbf(dist):
    """GaussianSmearing.forward, models/common.py:26-28 (20 fixed centres, coeff -0.5)."""
    off = torch.tensor(RBF_CENTRES, dtype=dist.dtype)
    d = dist.view(-1, 1) - off.view(1, -1)
    return torch.exp(-0.5 * d * d)


def mlp(sd, prefix, x):
    """MLP.forward, models/common.py:50-67: Linear -> LayerNorm(eps 1e-5) -> ReLU -> Linear."""
    y = F.linear(x, sd[prefix + '.net.0.weight'], sd[prefix + '.net.0.bias'])
    y = F.layer_norm(y, (y.shape[-1],), sd[prefix + '.net.1.weight'], sd[prefix + '.net.1.bias'], 1e-5)
    y = F.relu(y)
    return F.linear(y, sd[prefix + '.net.3.weight'], sd[prefix + '.net.3.bias'])


def seg_softmax(src, index, n):
    """torch_scatter.scatter_softmax(src, index, dim=0) restated: per-segment max shift / exp / sum /
    divide (call sites models/uni_transformer.py:77,:147)."""
    idx = index.view(-1, 1).expand_as(src)
    mx = torch.full((n, src.shape[1]), float('-inf'), dtype=src.dtype)
    mx = mx.scatter_reduce(0, idx, src, reduce='amax', include_self=True)
    ex = (src - mx[index]).exp()
    s = torch.zeros((n, src.shape[1]), dtype=src.dtype).index_add_(0, index, ex)
    return ex / s[index]


# --------------------------------------------------------------------------------------
# kNN graph  (models/uni_transformer.py:466-468 -> torch_geometric.nn.knn_graph -> torch_cluster.knn)
# --------------------------------------------------------------------------------------
def knn_dense(x, mol_ptr, k):
    """Dense neighbour table.  Returns nbr [N, k+1] int64 (global atom index, -1 padded) and deg [N].

    Semantics (torch_cluster 1.6.0 knn(x, x, k+1, batch, batch) followed by the `row != col` mask of
    knn_graph(loop=False)): per molecule, candidates sorted ascending by the key (d2, index) where
    d2 = fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)), dx = fl(x_i - x_j) in fp32 without FMA contraction;
    keep the first min(k+1, n); drop the entry whose index equals the centre.  Neighbours keep their
    ascending-distance order.
    """
    N = x.shape[0]
    nbr = torch.full((N, k + 1), -1, dtype=torch.long)
    deg = torch.zeros(N, dtype=torch.long)
    sizes = (mol_ptr[1:] - mol_ptr[:-1])
    for n in torch.unique(sizes).tolist():
        if n == 0:
            continue
        mols = torch.nonzero(sizes == n).flatten()
        starts = mol_ptr[:-1][mols]
        ar = torch.arange(n)
        gidx = starts[:, None] + ar[None, :]                     # [G, n]
        xs = x[gidx]                                             # [G, n, 3]
        d = xs[:, :, None, :] - xs[:, None, :, :]
        dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
        d2 = (dx * dx + dy * dy) + dz * dz                       # each op rounded separately
        order = torch.sort(d2, dim=2, stable=True).indices[:, :, :min(k + 1, n)]   # [G, n, kk]
        centre = ar.view(1, n, 1).expand_as(order)
        keep = order != centre
        # stable compaction of the kept entries to the left
        pos = torch.cumsum(keep.long(), dim=2) - 1
        glob = order + starts[:, None, None]
        out = torch.full((mols.numel(), n, k + 1), -1, dtype=torch.long)
        gi, ai, si = torch.nonzero(keep, as_tuple=True)
        out[gi, ai, pos[gi, ai, si]] = glob[gi, ai, si]
        nbr[gidx.flatten()] = out.view(-1, k + 1)
        deg[gidx.flatten()] = keep.sum(2).flatten()
    return nbr, deg


def knn_edges(x, mol_ptr, k):
    """edge_index [2, E] exactly as the reference sees it: row 0 = src (neighbour), row 1 = dst
    (centre); grouped by centre ascending, neighbours by ascending distance."""
    nbr, deg = knn_dense(x, mol_ptr, k)
    mask = nbr >= 0
    dst = torch.arange(x.shape[0])[:, None].expand_as(nbr)[mask]
    src = nbr[mask]
    return torch.stack([src, dst], 0)


# --------------------------------------------------------------------------------------
# the network  (models/uni_transformer.py, models/molopt_score_model.py:286-320)
# --------------------------------------------------------------------------------------
def invariant_shape_emb(sd, shape):
    """InvariantShapeEmbLayer.forward, models/uni_transformer.py:181-189.  shape [B,S,3] -> [B,S]."""
    m = shape.mean(dim=1)
    mn = m / ((m * m).sum(-1, keepdim=True) + EPS)
    inv = torch.einsum('bij,bj->bi', shape, mn)
    return mlp(sd, 'refine_net.invariant_shape_layer.hidden_layer', inv)


def edge_gate(sd, x, src, dst):
    """_pred_ew, models/uni_transformer.py:475-481."""
    dist = torch.norm(x[dst] - x[src], p=2, dim=-1, keepdim=True)
    return torch.sigmoid(mlp(sd, 'refine_net.edge_pred_layer', rbf(dist)))


def vn_linear_leaky_relu_nodes(sd, prefix, z, training, bn_stats_out=None, slope=0.2):
    """VNLinearLeakyReLU(dim=4).forward on z [N, Cin, 3] (models/shape_vn_layers.py:95-110) with
    VNBatchNorm (:49-61) -> nn.BatchNorm1d over the norms [N, Cout]."""
    p = torch.einsum('oc,ncd->nod', sd[prefix + '.map_to_feat.weight'], z)
    norm = torch.norm(p, dim=2) + EPS
    w, b = sd[prefix + '.batchnorm.bn.weight'], sd[prefix + '.batchnorm.bn.bias']
    if training:
        mean = norm.mean(0)
        var = norm.var(0, unbiased=False)
        if bn_stats_out is not None:
            bn_stats_out.append((mean, norm.var(0, unbiased=True) if norm.shape[0] > 1 else var))
    else:
        mean, var = sd[prefix + '.batchnorm.bn.running_mean'], sd[prefix + '.batchnorm.bn.running_var']
    norm_bn = (norm - mean) / torch.sqrt(var + 1e-5) * w + b
    p = p / norm.unsqueeze(2) * norm_bn.unsqueeze(2)
    d = torch.einsum('oc,ncd->nod', sd[prefix + '.map_to_dir.weight'], z)
    dot = (p * d).sum(2, keepdim=True)
    mask = (dot >= 0).float()
    dn = (d * d).sum(2, keepdim=True)
    return slope * p + (1 - slope) * (mask * p + (1 - mask) * (p - (dot / (dn + EPS)) * d))


def x2h_layer(sd, pre, h, r_feat, src, dst, inv_atoms, e_w, n_heads):
    """BaseX2HAttLayer.forward, models/uni_transformer.py:48-90 (shape_mode='attention')."""
    N = h.shape[0]
    kv = torch.cat([r_feat, h[dst], h[src], inv_atoms[dst]], -1)
    dh = h.shape[1] // n_heads
    k = mlp(sd, pre + '.hk_func', kv).view(-1, n_heads, dh)
    v = (mlp(sd, pre + '.hv_func', kv) * e_w.view(-1, 1)).view(-1, n_heads, dh)
    q = mlp(sd, pre + '.hq_func', h).view(-1, n_heads, dh)
    alpha = seg_softmax((q[dst] * k / np.sqrt(dh)).sum(-1), dst, N)
    m = alpha.unsqueeze(-1) * v
    out = torch.zeros((N, n_heads, dh), dtype=h.dtype).index_add_(0, dst, m).view(N, -1)
    out = mlp(sd, pre + '.node_output', torch.cat([out, h], -1))
    return out + h


def h2x_layer(sd, pre, h, x, rel_x, r_feat, src, dst, inv_atoms, shape_atoms, e_w, n_heads, training,
              bn_stats_out=None):
    """BaseH2XAttLayer.forward, models/uni_transformer.py:121-162 (shape_mode='attention_residue')."""
    N = h.shape[0]
    kv = torch.cat([r_feat, h[dst], h[src], inv_atoms[dst]], -1)
    dh = h.shape[1] // n_heads
    k = mlp(sd, pre + '.xk_func', kv).view(-1, n_heads, dh)
    v = mlp(sd, pre + '.xv_func', kv) * e_w.view(-1, 1)
    v = v.unsqueeze(-1) * rel_x.unsqueeze(1)
    q = mlp(sd, pre + '.xq_func', h).view(-1, n_heads, dh)
    alpha = seg_softmax((q[dst] * k / np.sqrt(dh)).sum(-1), dst, N)
    m = alpha.unsqueeze(-1) * v
    out = torch.zeros((N, n_heads, 3), dtype=h.dtype).index_add_(0, dst, m)
    z = torch.cat((x.unsqueeze(1), out, shape_atoms), dim=1)
    res = vn_linear_leaky_relu_nodes(sd, pre + '.shape_linear', z, training, bn_stats_out).mean(dim=1)
    return out.mean(dim=1) + res


def refine_net(sd, cfg, h, x, mol_ptr, shape, training, bn_stats_out=None, nbr_out=None):
    """UniTransformerO2TwoUpdateGeneral.forward, models/uni_transformer.py:483-540 with
    AttentionLayerO2TwoUpdateNodeGeneral.forward :289-333 inlined (num_x2h = num_h2x = 1,
    sync_twoup False, edge_feat_dim 0, topo_emb_type None)."""
    sizes = mol_ptr[1:] - mol_ptr[:-1]
    batch = torch.repeat_interleave(torch.arange(sizes.numel()), sizes)
    inv_atoms = invariant_shape_emb(sd, shape)[batch]
    shape_atoms = shape[batch]
    for _ in range(cfg['num_blocks']):
        edge_index = knn_edges(x, mol_ptr, cfg['knn'])
        if nbr_out is not None:
            nbr_out.append(edge_index)
        src, dst = edge_index
        e_w = edge_gate(sd, x, src, dst)
        for l in range(cfg['num_layers']):
            pre = 'refine_net.base_block.%d' % l
            rel_x = x[dst] - x[src]
            dist = torch.norm(rel_x, p=2, dim=-1, keepdim=True)
            r_feat = rbf(dist)      # outer_product(ones[E,1], r) == r   (models/common.py:70-77)
            h = x2h_layer(sd, pre + '.x2h_layers.0', h, r_feat, src, dst, inv_atoms, e_w, cfg['n_heads'])
            dx = h2x_layer(sd, pre + '.h2x_layers.0', h, x, rel_x, r_feat, src, dst, inv_atoms, shape_atoms,
                           e_w, cfg['n_heads'], training, bn_stats_out)
            x = x + dx
    return x, h


def time_embedding(sd, t, dim):
    """SinusoidalPosEmb + time_emb, models/molopt_score_model.py:154-166,247-252."""
    half = dim // 2
    e = np.log(10000) / (half - 1)
    w = torch.exp(torch.arange(half) * -e)
    emb = t[:, None] * w[None, :]
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    y = F.linear(emb, sd['time_emb.1.weight'], sd['time_emb.1.bias'])
    y = F.silu(y)
    return F.linear(y, sd['time_emb.3.weight'], sd['time_emb.3.bias'])


def forward(sd, cfg, pos, v, mol_ptr, shape, t, training=True, bn_stats_out=None, nbr_out=None):
    """ScorePosNet3D.forward, models/molopt_score_model.py:286-320.
    pos [N,3] f32, v [N] i64, mol_ptr [B+1] i64, shape [B,S,3], t [B] i64.
    Returns (pred_pos [N,3], pred_h [N,H], pred_v logits [N,C])."""
    sizes = mol_ptr[1:] - mol_ptr[:-1]
    batch = torch.repeat_interleave(torch.arange(sizes.numel()), sizes)
    C = sd['v_inference.2.weight'].shape[0]
    onehot = F.one_hot(v, C).float()
    tf = time_embedding(sd, t, cfg['time_emb_dim'])[batch]
    h0 = F.linear(torch.cat([onehot, tf], -1), sd['ligand_atom_emb.weight'], sd['ligand_atom_emb.bias'])
    x, h = refine_net(sd, cfg, h0, pos, mol_ptr, shape, training, bn_stats_out, nbr_out)
    y = F.linear(h, sd['v_inference.0.weight'], sd['v_inference.0.bias'])
    y = F.softplus(y) - math.log(2.0)           # ShiftedSoftplus, models/common.py:39-45
    logits = F.linear(y, sd['v_inference.2.weight'], sd['v_inference.2.bias'])
    return x, h, logits


# --------------------------------------------------------------------------------------
# posterior / categorical  (models/molopt_score_model.py:64-68,98-113,323-404,658-669)
# --------------------------------------------------------------------------------------
def log_add_exp(a, b):
    m = torch.max(a, b)
    return m + torch.log(torch.exp(a - m) + torch.exp(b - m))


def posterior_step(tabs, x0, logits, x_t, v_t, t_atoms, noise_pos, noise_u):
    """One reverse step of sample_diffusion's default branch (:658-669) with injected noise.
    t_atoms [N] i64.  Returns (x_next [N,3], v_next [N] i64, log_v0 [N,C], log_post [N,C])."""
    C = logits.shape[1]
    c0 = tabs['posterior_mean_c0_coef'][t_atoms].unsqueeze(-1)
    ct = tabs['posterior_mean_ct_coef'][t_atoms].unsqueeze(-1)
    mean = c0 * x0 + ct * x_t
    logvar = tabs['posterior_logvar'][t_atoms].unsqueeze(-1)
    nz = (1 - (t_atoms == 0).float()).unsqueeze(-1)
    x_next = mean + nz * (0.5 * logvar).exp() * noise_pos
    log_v0 = F.log_softmax(logits, dim=-1)
    log_vt = torch.log(F.one_hot(v_t, C).float().clamp(min=1e-30))
    tm1 = torch.where(t_atoms - 1 < 0, torch.zeros_like(t_atoms), t_atoms - 1)
    lnC = np.log(C)
    a = log_add_exp(log_v0 + tabs['log_alphas_cumprod_v'][tm1].unsqueeze(-1),
                    tabs['log_one_minus_alphas_cumprod_v'][tm1].unsqueeze(-1) - lnC)
    b = log_add_exp(log_vt + tabs['log_alphas_v'][t_atoms].unsqueeze(-1),
                    tabs['log_one_minus_alphas_v'][t_atoms].unsqueeze(-1) - lnC)
    un = a + b
    post = un - torch.logsumexp(un, dim=-1, keepdim=True)
    gumbel = -torch.log(-torch.log(noise_u + 1e-30) + 1e-30)
    v_next = (gumbel + post).argmax(dim=-1)
    return x_next, v_next, log_v0, post


def sample(sd, cfg, tabs, pos, v, mol_ptr, shape, t_begin, num_steps, noise_fn, training=True,
           keep_traj=False):
    """sample_diffusion default branch (models/molopt_score_model.py:533-697), noise injected by
    noise_fn(step_index) -> (randn [N,3], rand [N,C]) in the reference's draw order."""
    sizes = mol_ptr[1:] - mol_ptr[:-1]
    B = sizes.numel()
    batch = torch.repeat_interleave(torch.arange(B), sizes)
    traj = []
    for s, i in enumerate(range(t_begin, t_begin - num_steps, -1)):
        t = torch.full((B,), i, dtype=torch.long)
        x0, _, logits = forward(sd, cfg, pos, v, mol_ptr, shape, t, training)
        eps, u = noise_fn(s)
        pos, v, lv0, post = posterior_step(tabs, x0, logits, pos, v, t[batch], eps, u)
        if keep_traj:
            traj.append((x0, logits, pos, v, lv0, post))
    return pos, v, traj


# --------------------------------------------------------------------------------------
# VN-DGCNN shape encoder (models/shape_pointcloud_modelAE.py:207-255, models/shape_vn_layers.py)
# --------------------------------------------------------------------------------------
def enc_knn(feat, k):
    """knn(), models/shape_vn_layers.py:286-292.  feat [B, D, P] -> idx [B, P, k] (includes self)."""
    inner = -2 * torch.matmul(feat.transpose(2, 1), feat)
    xx = torch.sum(feat ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    return pd.topk(k=k, dim=-1)[1]


def enc_graph_feature(x, k):
    """get_graph_feature_cross(if_cross=False), models/shape_vn_layers.py:257-284.
    x [B, C, 3, P] -> [B, 2C, 3, P, k] = cat(f_j - f_i, f_i)."""
    B, C, _, P = x.shape
    flat = x.reshape(B, C * 3, P)
    idx = enc_knn(flat, k)                                           # [B,P,k]
    xt = flat.transpose(2, 1).contiguous()                           # [B,P,3C]
    nb = torch.gather(xt.unsqueeze(1).expand(B, P, P, C * 3), 2,
                      idx.unsqueeze(-1).expand(B, P, k, C * 3))      # [B,P,k,3C]
    nb = nb.view(B, P, k, C, 3)
    ctr = xt.view(B, P, 1, C, 3).expand(B, P, k, C, 3)
    return torch.cat((nb - ctr, ctr), dim=3).permute(0, 3, 4, 1, 2).contiguous()


def enc_vn_layer(w_feat, w_dir, bn_w, bn_b, x, slope=0.2, running=None):
    """VNLinearLeakyReLU (dim 5 or 4), models/shape_vn_layers.py:95-110; BatchNorm over every axis
    except channel with batch statistics (running=None) or running stats (mean, var)."""
    p = torch.einsum('oc,bc...->bo...', w_feat, x)
    norm = torch.norm(p, dim=2) + EPS                                 # [B,O,P(,k)]
    red = [d for d in range(norm.dim()) if d != 1]
    if running is None:
        mean = norm.mean(dim=red, keepdim=True)
        var = norm.var(dim=red, unbiased=False, keepdim=True)
    else:
        shp = [1, -1] + [1] * (norm.dim() - 2)
        mean, var = running[0].view(shp), running[1].view(shp)
    shp = [1, -1] + [1] * (norm.dim() - 2)
    nbn = (norm - mean) / torch.sqrt(var + 1e-5) * bn_w.view(shp) + bn_b.view(shp)
    p = p / norm.unsqueeze(2) * nbn.unsqueeze(2)
    d = torch.einsum('oc,bc...->bo...', w_dir, x)
    dot = (p * d).sum(2, keepdim=True)
    mask = (dot >= 0).float()
    dn = (d * d).sum(2, keepdim=True)
    return slope * p + (1 - slope) * (mask * p + (1 - mask) * (p - (dot / (dn + EPS)) * d))


def encoder_forward(w, clouds, k=20, training=True):
    """VN_DGCNN_Encoder.forward, models/shape_pointcloud_modelAE.py:231-255.
    w: dict with conv_pos.*, conv_c.* (state_dict names) and blocks.{i}.map_to_feat/map_to_dir/
    batchnorm.bn.{weight,bias} taken from the live module (the blocks are unregistered and ALWAYS
    use batch statistics).  clouds [B,1,P,3] -> latent [B,L,3]."""
    x = clouds.transpose(2, 3)
    run = None if training else (w['conv_pos.batchnorm.bn.running_mean'], w['conv_pos.batchnorm.bn.running_var'])
    hid = enc_vn_layer(w['conv_pos.map_to_feat.weight'], w['conv_pos.map_to_dir.weight'],
                       w['conv_pos.batchnorm.bn.weight'], w['conv_pos.batchnorm.bn.bias'],
                       enc_graph_feature(x, k), running=run).mean(dim=-1)
    hs = []
    i = 0
    while ('blocks.%d.map_to_feat.weight' % i) in w:
        p = 'blocks.%d.' % i
        hid = enc_vn_layer(w[p + 'map_to_feat.weight'], w[p + 'map_to_dir.weight'],
                           w[p + 'batchnorm.bn.weight'], w[p + 'batchnorm.bn.bias'],
                           enc_graph_feature(hid, k)).mean(dim=-1)
        hs.append(hid)
        i += 1
    cat = torch.cat(hs, dim=1)
    run = None if training else (w['conv_c.batchnorm.bn.running_mean'], w['conv_c.batchnorm.bn.running_var'])
    lat = enc_vn_layer(w['conv_c.map_to_feat.weight'], w['conv_c.map_to_dir.weight'],
                       w['conv_c.batchnorm.bn.weight'], w['conv_c.batchnorm.bn.bias'], cat, running=run)
    return lat.mean(dim=-1)


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench
# --------------------------------------------------------------------------------------
DEFAULT_CFG = dict(num_blocks=1, num_layers=8, hidden_dim=128, n_heads=16, knn=32, num_r_gaussian=20,
                   shape_dim=32, shape_latent_dim=32, time_emb_dim=8, num_diffusion_timesteps=1000,
                   schedule_pos=dict(beta_schedule='sigmoid', beta_start=1e-7, beta_end=0.01, s=6),
                   schedule_v=dict(beta_schedule='cosine', s=0.01))


def mol_ptr_from_sizes(sizes):
    return torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(torch.as_tensor(sizes, dtype=torch.long), 0)])


# ---------------------------------------------------------------------------------------------
# Point-cloud shape guidance: models/molopt_score_model.py:699-740, applied to the predicted x0 while
# t > grad_step (:582-591).  Every atom is independent.  The reference draws its pull-back scalars from numpy's global
# RNG, one per atom that is still "far" in iteration j, in ascending atom order; here they are an explicit dense input
# u[j][atom] (float64; entries of atoms that are not far in iteration j are never read).
# Arithmetic follows the reference's numpy / sklearn float64 evaluation order exactly (KDTree.query = exact Euclidean
# distances in double, sum over x, y, z in that order; np.mean over the 3 neighbours = ((a + b) + c) / 3).
# ---------------------------------------------------------------------------------------------
def _three_nn(points, cloud):
    """points [n,3] f64, cloud [M,3] f64 -> (dists [n,3] ascending, idx [n,3]); ties broken by lower index."""
    dx = points[:, None, 0] - cloud[None, :, 0]
    dy = points[:, None, 1] - cloud[None, :, 1]
    dz = points[:, None, 2] - cloud[None, :, 2]
    d2 = (dx * dx + dy * dy) + dz * dz
    order = torch.sort(d2, dim=1, stable=True)
    return torch.sqrt(order.values[:, :3]), order.indices[:, :3]


def pointcloud_guidance(pos, cloud, radius, u, ratio=0.2, max_iter=5):
    """pos [N,3] f32 (predicted x0), cloud [M,3] f64, u [max_iter,N] f64 in [0,1) -> guided pos [N,3] f32."""
    cloud = cloud.to(torch.float64)
    p = pos.to(torch.float64).clone()
    out = pos.clone()
    span = 0.8 - ratio
    dists, idx = _three_nn(p, cloud)
    far = ((dists[:, 0] + dists[:, 1]) + dists[:, 2]) / 3.0 > radius
    atoms = torch.nonzero(far).flatten()
    pts, nn = p[atoms], idx[atoms]
    j = 0
    while atoms.numel() > 0 and j < max_iter:
        c = cloud[nn]                                              # [n,3 neighbours,3]
        near = ((c[:, 0] + c[:, 1]) + c[:, 2]) / 3.0
        s = (u[j, atoms] * span + ratio)[:, None]
        new = pts - s * (pts - near)
        dists, idx = _three_nn(new, cloud)
        inside = ((dists[:, 0] + dists[:, 1]) + dists[:, 2]) / 3.0 < radius
        out[atoms[inside]] = new[inside].to(torch.float32)
        atoms, pts, nn = atoms[~inside], new[~inside], idx[~inside]
        j += 1
    if j == max_iter and atoms.numel() > 0:
        out[atoms] = pts.to(torch.float32)
    return out


# ---------------------------------------------------------------------------------------------
# Alignment-free Gaussian-overlap shape Tanimoto: utils/evaluation/shaep_utils.py:59-83 (get_ROCS, all
# atoms alpha = 0.81, prefactor 0.8).  Float64, explicit pair distances.
# ---------------------------------------------------------------------------------------------
def rocs_constants(prefactor=0.8, alpha=0.81):
    """The reference builds its per-atom constants with torch.ones(...) * x, i.e. in float32, and only the pair distances
    are float64: (k, coef, den) with  term = coef * exp(-k R^2) / den  (shaep_utils.py:59-66)."""
    a, p = np.float32(alpha), np.float32(prefactor)
    k = np.float32(np.float32(a * a) / np.float32(a + a))
    coef = np.float32(np.float32(math.pi ** 1.5) * np.float32(p * p))
    den = np.float32(np.float32(a + a) ** np.float32(1.5))
    return float(k), float(coef), float(den)


def _vab(c1, c2, consts):
    k, coef, den = consts
    d = c1[:, None, :] - c2[None, :, :]
    r2 = torch.sqrt((d * d).sum(-1)) ** 2.0          # the reference squares torch.cdist
    return (coef * torch.exp(-k * r2) / den).sum()


def get_rocs(centers_1, centers_2, prefactor=0.8, alpha=0.81):
    c1, c2 = centers_1.to(torch.float64), centers_2.to(torch.float64)
    cs = rocs_constants(prefactor, alpha)
    vaa, vbb, vab = _vab(c1, c1, cs), _vab(c2, c2, cs), _vab(c1, c2, cs)
    return vab / (vaa + vbb - vab)


def get_rocs_batch(pos, mol_ptr, ref, ref_ptr):
    """Per-molecule Tanimoto of generated centres pos[mol_ptr[m]:mol_ptr[m+1]] against ref[ref_ptr[m]:ref_ptr[m+1]]."""
    return torch.stack([get_rocs(pos[int(mol_ptr[m]):int(mol_ptr[m + 1])], ref[int(ref_ptr[m]):int(ref_ptr[m + 1])])
                        for m in range(len(mol_ptr) - 1)])


# ---------------------------------------------------------------------------------------------
# Stability check: utils/evaluation/analyze.py:249-297 (get_bond_order, check_stability).  Table driven: thr
# [3,E,E] int (bond length + margin in pm, shapemol_b200/chem_tables.py), allowed [E].  Distances in the positions' dtype
# (float32 from the sampler), every operation individually rounded, then x 100 -- numpy >= 2 scalar promotion keeps float32.
# ---------------------------------------------------------------------------------------------
def check_stability(positions, elem, thr, allowed, hs=False):
    """positions [n,3] f32, elem [n] element index -> (molecule_stable, nr_stable_atoms, n, nr_bonds [n])."""
    p = positions.to(torch.float32)
    n = p.shape[0]
    dx, dy, dz = p[:, None, 0] - p[None, :, 0], p[:, None, 1] - p[None, :, 1], p[:, None, 2] - p[None, :, 2]
    d = torch.sqrt((dx * dx + dy * dy) + dz * dz) * 100.0
    e = elem.long()
    t1, t2, t3 = (thr[k][e[:, None], e[None, :]].to(torch.float32) for k in range(3))
    o1 = d < t1
    o2 = o1 & (d < t2)
    o3 = o2 & (d < t3)
    order = o1.long() + o2.long() + o3.long()
    order.fill_diagonal_(0)
    nr = order.sum(1)
    al = allowed[e].long()
    ok = (al == nr) if hs else ((al >= nr) & (nr > 0))
    return bool(ok.all()) if n else True, int(ok.sum()), n, nr
