"""Synthetic inputs of SURVEY.md §8(d): seeded tile graphs, slide data for the builder, and deterministic weights.
Everything is generated on the CPU with torch/numpy generators so that the oracle and the CUDA path see identical
inputs; nothing here is timed."""
import math

import numpy as np
import torch

from .graph import Batch, Data


def tile_graph(num_nodes, node_features=768, seed=1234, avg_degree=16, device="cpu", dtype=torch.float32):
    """Tile graph: nodes on a ceil(sqrt(N))-wide jittered grid (row-major), undirected pairs = right / down /
    down-right / down-left grid neighbours (~4N) + 4N uniformly random pairs, both directions emitted as
    (s,t),(t,s) like the builder does -> E ~= 16N. pos = normalised jittered coordinates in [0,1)^2."""
    g = torch.Generator().manual_seed(seed)
    n = int(num_nodes)
    w = int(math.ceil(math.sqrt(n)))
    hgt = (n + w - 1) // w
    idx = torch.arange(n)
    r, c = idx // w, idx % w
    pairs = []
    for dr, dc in ((0, 1), (1, 0), (1, 1), (1, -1)):
        rr, cc = r + dr, c + dc
        ok = (rr < hgt) & (cc >= 0) & (cc < w)
        j = rr * w + cc
        ok &= j < n
        pairs.append(torch.stack([idx[ok], j[ok]]))
    n_rand = max(0, (avg_degree // 2 - 4)) * n
    if n_rand > 0:
        pairs.append(torch.randint(0, n, (2, n_rand), generator=g))
    und = torch.cat(pairs, dim=1)
    edge_index = torch.stack([und, und.flip(0)], dim=2).reshape(2, -1).contiguous()     # (s,t),(t,s) interleaved
    jit = (torch.rand(n, 2, generator=g) - 0.5) * 0.3
    pos = torch.stack([(c.float() + 0.5 + jit[:, 0]) / w, (r.float() + 0.5 + jit[:, 1]) / max(hgt, 1)], dim=1)
    x = torch.randn(n, node_features, generator=g, dtype=dtype)
    d = Data(x=x, edge_index=edge_index.long(), pos=pos.float())
    return d.to(device) if device != "cpu" else d


def multimag_graph(num_nodes, node_features=1024, seed=1234, device="cpu"):
    """one multi-magnification slide graph. Three nested grids (5x / 20x / 40x, node populations 1:16:64:
    g x g, 4g x 4g, 8g x 8g cells over the same unit square), links = the 8-neighbourhood inside each level (4 undirected pairs per
    node), one parent/child link per 20x / 40x node to the coarser cell that contains it, and 16 "morphological" links per node
    (8 undirected pairs: 4 inside a +-8-cell window of the same level, 4 uniformly random over the whole slide). Both directions are
    emitted as (s,t),(t,s). Nodes are ordered by the coarse-grid row band that contains them (then level, then row-major), so that
    a contiguous 1-D node partition is a partition into horizontal bands of the slide. Returns Data with x, edge_index, pos and
    `level` (0 / 1 / 2 per node)."""
    gen = torch.Generator().manual_seed(seed)
    n_target = int(num_nodes)
    g = max(1, int(round(math.sqrt(n_target / 81.0))))
    sizes = [g, 4 * g, 8 * g]
    lv, gx, gy = [], [], []
    for l, w in enumerate(sizes):
        i = torch.arange(w * w)
        lv.append(torch.full((w * w,), l, dtype=torch.long)); gx.append(i % w); gy.append(i // w)
    lv, gx, gy = torch.cat(lv), torch.cat(gx), torch.cat(gy)
    wl = torch.tensor(sizes)[lv]
    band = (gy * g) // wl                                             # coarse row band of the cell
    key = (band * 3 + lv) * (64 * g * g) + gy * wl + gx
    order = torch.argsort(key, stable=True)
    n_all = order.numel()
    keep = order[:min(n_target, n_all)] if n_all >= n_target else order
    newid = torch.full((n_all,), -1, dtype=torch.long)
    newid[keep] = torch.arange(keep.numel())
    off = [0, sizes[0] ** 2, sizes[0] ** 2 + sizes[1] ** 2]

    def cell(l, x, y):
        return off[l] + y * sizes[l] + x
    pairs = []
    for l, w in enumerate(sizes):
        i = torch.arange(w * w)
        x, y = i % w, i // w
        for dy, dx in ((0, 1), (1, 0), (1, 1), (1, -1)):
            ok = (y + dy < w) & (x + dx >= 0) & (x + dx < w)
            pairs.append(torch.stack([cell(l, x[ok], y[ok]), cell(l, x[ok] + dx, y[ok] + dy)]))
        if l > 0:
            f = sizes[l] // sizes[l - 1]
            pairs.append(torch.stack([cell(l, x, y), cell(l - 1, x // f, y // f)]))                      # child -> parent
        for _ in range(4):                                                                            # local morphological links
            ddx = torch.randint(-8, 9, (w * w,), generator=gen); ddy = torch.randint(-8, 9, (w * w,), generator=gen)
            x2, y2 = (x + ddx).clamp(0, w - 1), (y + ddy).clamp(0, w - 1)
            pairs.append(torch.stack([cell(l, x, y), cell(l, x2, y2)]))
    und = torch.cat(pairs, dim=1)
    und = torch.cat([und, torch.stack([torch.arange(n_all).repeat(4), torch.randint(0, n_all, (4 * n_all,), generator=gen)])], dim=1)
    und = newid[und]
    und = und[:, (und >= 0).all(dim=0)]
    edge_index = torch.stack([und, und.flip(0)], dim=2).reshape(2, -1).contiguous()
    n = keep.numel()
    jit = (torch.rand(n, 2, generator=gen) - 0.5) * 0.3
    wk = wl[keep].float()
    pos = torch.stack([(gx[keep].float() + 0.5 + jit[:, 0]) / wk, (gy[keep].float() + 0.5 + jit[:, 1]) / wk], dim=1)
    x = torch.randn(n, node_features, generator=gen)
    d = Data(x=x, edge_index=edge_index.long(), pos=pos.float(), level=lv[keep])
    return d.to(device) if device != "cpu" else d


def edge_cut_fraction(edge_index, num_nodes, world):
    """Fraction of directed edges whose endpoints fall into different blocks of the even contiguous 1-D node partition."""
    rows = -(-int(num_nodes) // int(world))
    if edge_index.numel() == 0 or world <= 1:
        return 0.0
    return float(((edge_index[0] // rows) != (edge_index[1] // rows)).float().mean().item())


def tile_batch(num_graphs, nodes_per_graph, node_features=768, seed=1234, device="cpu"):
    b = Batch.from_data_list([tile_graph(nodes_per_graph, node_features, seed + i) for i in range(num_graphs)])
    return b.to(device) if device != "cpu" else b


def seeded_state_dict(model, seed=0, perturb=True):
    """Deterministic weights keyed on parameter NAME (independent of construction order, so the reference model
    and this package get identical tensors): xavier-uniform matrices; with perturb=True biases are U(-0.1,0.1) and
    norm scales 1+U(-0.1,0.1) so that every epilogue term is exercised."""
    sd = {}
    for name, t in model.state_dict().items():
        h = 0
        for ch in name:
            h = (h * 131 + ord(ch)) % 2147483629
        g = torch.Generator().manual_seed((h + 7919 * seed) % 2147483629)
        if not t.dtype.is_floating_point:
            sd[name] = t.clone()
            continue
        shape = tuple(t.shape)
        if name.endswith("running_var"):
            v = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif name.endswith("running_mean"):
            v = 0.1 * (torch.rand(shape, generator=g) - 0.5)
        elif t.dim() >= 2 and "pos_encoding" not in name and "global_token" not in name:
            fan_out, fan_in = shape[0], int(np.prod(shape[1:]))
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif "global_token" in name:
            v = torch.randn(shape, generator=g)
        elif "pos_encoding" in name:
            v = torch.randn(shape, generator=g) * 0.02
        elif name.endswith("weight"):          # 1-D: norm scale
            v = 1.0 + (0.2 * (torch.rand(shape, generator=g) - 0.5) if perturb else 0.0) * torch.ones(shape)
        else:                                   # bias
            v = 0.2 * (torch.rand(shape, generator=g) - 0.5) if perturb else torch.zeros(shape)
        sd[name] = v.to(t.dtype)
    return sd


def clustered_features(n, d=768, num_clusters=64, seed=7, sigma_lo=0.018, sigma_hi=0.028):
    """Mixture of Gaussian clusters with unit-norm centres; per-cluster sigma ~ U[lo,hi] puts the within-cluster cosine
    around the 0.7 edge threshold."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((num_clusters, d))
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    sig = rng.uniform(sigma_lo, sigma_hi, size=num_clusters)
    which = rng.integers(0, num_clusters, size=n)
    x = centres[which] + rng.standard_normal((n, d)) * sig[which][:, None]
    return x.astype(np.float32)


def jittered_grid_coords(n, seed=7, extent=100000):
    rng = np.random.default_rng(seed + 1)
    w = int(math.ceil(math.sqrt(n)))
    pitch = extent // (w + 1)
    i = np.arange(n)
    x = ((i % w) + 0.5 + rng.uniform(-0.15, 0.15, n)) * pitch
    y = ((i // w) + 0.5 + rng.uniform(-0.15, 0.15, n)) * pitch
    return x.astype(np.int64), y.astype(np.int64), (extent, extent)
