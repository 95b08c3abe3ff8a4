"""Test-infrastructure shim: torch_geometric.nn.knn_graph restated from the published
semantics of torch-cluster 1.6.0 (knn(x, x, k+1) within each batch segment, then drop self
by index; flow='source_to_target' => edge_index[0] = neighbour, edge_index[1] = centre).

Distance is the canonical fp32 direct-difference form fixed in DESIGN.md:
d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)); ordering key (d2, index) ascending (the CUDA
kernel of torch_cluster keeps the lower index on ties)."""
import torch


def knn_graph(x, k, batch=None, loop=False, flow='source_to_target', cosine=False, num_workers=1):
    assert flow == 'source_to_target' and not cosine
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long, device=x.device)
    rows, cols = [], []
    kk = k if loop else k + 1
    counts = torch.bincount(batch)
    start = 0
    for c in counts.tolist():
        if c == 0:
            continue
        xs = x[start:start + c]
        d = xs[:, None, :] - xs[None, :, :]
        dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        order = torch.sort(d2, dim=1, stable=True).indices[:, :min(kk, c)]   # [c, kk]
        centre = torch.arange(c, device=x.device)[:, None].expand_as(order)
        keep = order != centre if not loop else torch.ones_like(order, dtype=torch.bool)
        rows.append(centre[keep] + start)
        cols.append(order[keep] + start)
        start += c
    row = torch.cat(rows)
    col = torch.cat(cols)
    return torch.stack([col, row], dim=0)


def radius_graph(*args, **kwargs):
    raise NotImplementedError('radius_graph is not on the hot path')
