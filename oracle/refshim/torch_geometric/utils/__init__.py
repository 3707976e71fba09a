import torch


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, torch.stack([loops, loops])], dim=1), edge_attr


def degree(index, num_nodes=None, dtype=None):
    n = int(num_nodes) if num_nodes is not None else int(index.max()) + 1
    out = torch.zeros(n, dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.index_add_(0, index, torch.ones(index.numel(), dtype=out.dtype, device=index.device))


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    n = int(num_nodes) if num_nodes is not None else int(index.max()) + 1
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    smax = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    smax = smax.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    out = (src - smax.gather(0, idx)).exp()
    ssum = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, out)
    return out / (ssum.gather(0, idx) + 1e-16)
