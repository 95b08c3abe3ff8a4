"""Test-infrastructure shim: minimal pure-torch torch_scatter (2.0.9 semantics) so that the
UNMODIFIED reference modules under /root/reference import in the build container.
Used only by tests/golden/make_golden.py (fixture generation); never by the product path."""
import torch


def _bcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _bcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    idx_dim = dim if dim >= 0 else dim + src.dim()
    if index.dim() <= idx_dim:
        idx_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, idx_dim, None, dim_size)
    count[count < 1] = 1
    count = _bcast(count, out, dim)
    return out / count


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    index_b = _bcast(index, src, dim)
    size = list(src.size())
    size[dim] = dim_size if dim_size is not None else int(index.max()) + 1
    out = torch.full(size, float('-inf'), dtype=src.dtype, device=src.device)
    out = out.scatter_reduce(dim, index_b, src, reduce='amax', include_self=True)
    return out, None


def scatter_softmax(src, index, dim=-1, eps=1e-12, dim_size=None):
    index_b = _bcast(index, src, dim)
    mx, _ = scatter_max(src, index, dim, dim_size=dim_size)
    rec = src - mx.gather(dim, index_b)
    ex = rec.exp()
    s = scatter_sum(ex, index_b, dim, dim_size=dim_size)
    return ex / s.gather(dim, index_b)
