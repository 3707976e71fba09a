"""Host-side mirror of the reference's `core` layers (core/graph_layers.py, core/attention.py, core/diffusion.py)
and encoders (models/encoders.py): same constructor signatures, attribute names and state_dict keys; the
forward of every class dispatches to the sm_100a kernels in csrc/ through ops.py.

Inference semantics: dropout is the identity (eval mode) and no autograd graph is recorded — the backward
kernels are the "next" row of SURVEY.md §8(f).
"""
import math
import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .graph import GraphContext

class _IdCache:
    """Cache keyed on the identity of a Parameter (tensors cannot be dict keys by value); entries die with the parameter."""

    def __init__(self):
        self._d = {}

    def get(self, obj):
        return self._d.get(id(obj))

    def put(self, obj, value):
        i = id(obj)
        if i not in self._d:
            weakref.finalize(obj, self._d.pop, i, None)
        self._d[i] = value


_SPLIT_CACHE = _IdCache()
_CAT_CACHE = _IdCache()


def invalidate_weight_cache():
    """Drop every cached fp16 weight image. The caches re-make an image when a Parameter's storage address or `_version`
    changes (optimizer steps, load_state_dict, .to()); an in-place write through `param.data` changes neither and must be
    followed by this call."""
    for c in (_SPLIT_CACHE, _CAT_CACHE, _COLSPLIT_CACHE, _SCALAR_CACHE):
        c._d.clear()


def split_weight(param):
    """fp16 hi/lo planes of a weight Parameter, re-made when the parameter's storage or version changes."""
    key = (param.data_ptr(), param._version, tuple(param.shape))
    hit = _SPLIT_CACHE.get(param)
    if hit is None or hit[0] != key:
        hit = (key, ops.SplitWeight(param.detach()))
        _SPLIT_CACHE.put(param, hit)
    return hit[1]


def split_weight_cat(params, tag):
    """Planes of torch.cat(params) (e.g. q|k|v projection). The entry lives in a per-tag dict attached to the first part
    (dies with it) and is valid only while every part is the SAME object at the same address / version."""
    key = tuple((p.data_ptr(), p._version) for p in params)
    per_tag = _CAT_CACHE.get(params[0])
    if per_tag is None:
        per_tag = {}
        _CAT_CACHE.put(params[0], per_tag)
    hit = per_tag.get(tag)
    if hit is None or hit[0] != key or any(r() is not p for r, p in zip(hit[2], params)):
        w = torch.cat([p.detach() for p in params], dim=0)
        hit = (key, ops.SplitWeight(w), tuple(weakref.ref(p) for p in params))
        per_tag[tag] = hit
    return hit[1]


def _cat_bias(linears):
    if linears[0].bias is None:
        return None
    return torch.cat([l.bias.detach() for l in linears])


_UNIT_LN = {}


def _ln(m):
    """(gamma, beta) of the row-wise normalisation fused into a GEMM epilogue. InstanceNorm1d on the reference's 2-D [N, C] inputs
    normalises every ROW over its C entries without affine parameters (torch reads [N, C] as an unbatched (channels, length)
    signal): LayerNorm with gamma = 1, beta = 0 (same eps 1e-5, biased variance)."""
    if isinstance(m, nn.LayerNorm):
        return (m.weight.detach(), m.bias.detach())
    if isinstance(m, nn.InstanceNorm1d):
        dev = torch.device("cuda", torch.cuda.current_device())      # the module has no parameters to take the device from
        key = (m.num_features, dev)
        if key not in _UNIT_LN:
            _UNIT_LN[key] = (torch.ones(m.num_features, device=dev), torch.zeros(m.num_features, device=dev))
        return _UNIT_LN[key]
    return None


def _norm_act_torch(h, norm, act_name, training):
    """BatchNorm1d (+ activation) — a non-benchmarked constructor option (SURVEY 8(b) "non-default options"): the column-wise
    affine of eval-mode batch norm does not fit the row-wise GEMM epilogues, so these two element-wise steps run as plain torch
    CUDA ops on the GEMM output."""
    h = F.batch_norm(h, norm.running_mean, norm.running_var, norm.weight, norm.bias, training and norm.track_running_stats is not False,
                     norm.momentum if norm.momentum is not None else 0.0, norm.eps)
    return {"gelu": F.gelu, "relu": F.relu, "elu": F.elu}[act_name](h) if act_name else h


def _require_cuda(x):
    if not x.is_cuda:
        raise ops.KernelError("dgdm_histopath_lab_b200 runs on a B200 (CUDA) device only; got a CPU tensor. "
                              "There is no CPU fallback.")


def _ctx_for(x, edge_index, ctx):
    if ctx is not None:
        return ctx
    return GraphContext(edge_index, None, x.size(0))


# ------------------------------------------------------------------------------------------ graph layers
class GraphConvolution(nn.Module):
    """core/graph_layers.py:19-110. out = D^-1/2 (A+I) D^-1/2 (x W^T + edge_attr W_e^T) + b."""

    def __init__(self, in_channels, out_channels, edge_dim=None, bias=True, add_self_loops=True, normalize=True, **kwargs):
        super().__init__()
        if kwargs.get("aggr", "add") != "add":
            raise NotImplementedError("only aggr='add' (the reference default) is implemented")
        if not (add_self_loops and normalize):
            raise NotImplementedError("only normalize=True, add_self_loops=True (the reference defaults) are implemented")
        self.in_channels, self.out_channels, self.edge_dim = in_channels, out_channels, edge_dim
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.node_lin = nn.Linear(in_channels, out_channels, bias=False)
        self.edge_lin = nn.Linear(edge_dim, out_channels, bias=False) if edge_dim is not None else None
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.node_lin.weight)
        if self.edge_lin is not None:
            nn.init.xavier_uniform_(self.edge_lin.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    @torch.no_grad()
    def forward(self, x, edge_index=None, edge_attr=None, size=None, *, ctx=None, agg_attr=None, act=None, n_rows=None, part=None):
        """part (parallel.Part): x holds the rows this rank owns; the projected rows are all-gathered (the halo exchange) and
        the propagate runs on the owned destination rows only. agg_attr is then the local slice as well."""
        _require_cuda(x)
        ctx = _ctx_for(x, edge_index, ctx)
        csr = ctx.csr if isinstance(ctx, GraphContext) else ctx
        rows, h = None, None
        if part is not None:
            rows, n_rows = part.rows, None
            fused = getattr(part.comm, "fused_linear_gather", None)
            if fused is not None and self.out_channels <= 512:
                # GEMM -> all-gather in one kernel: the epilogue stores each tile into every rank's copy over NVLink
                w = split_weight(self.node_lin.weight)
                h = fused(lambda out, peers, mc: ops.linear(x, w, out=out, peers=peers, multicast=mc), part, self.out_channels)
            if h is None:
                h = part.gather(ops.linear(x, split_weight(self.node_lin.weight)))
        else:
            h = ops.linear(x, split_weight(self.node_lin.weight))
        addend = None
        if self.edge_lin is not None and (edge_attr is not None or agg_attr is not None):
            if agg_attr is None:
                agg_attr = ops.edge_attr_aggregate(csr, _pad_cols(edge_attr, self.edge_dim))
            # sum_e n_e W_e a_e = W_e (sum_e n_e a_e): one [N,32] GEMM instead of an [E+N, C] intermediate
            addend = ops.linear(agg_attr, split_weight(self.edge_lin.weight))
        return ops.spmm_gcn(csr, h, n_rows=n_rows, rows=rows, addend=addend,
                            bias=self.bias.detach() if self.bias is not None else None, act=act)


def _pad_cols(t, width):
    if t.size(1) == width:
        return t.contiguous()
    if t.size(1) > width:
        raise ops.KernelError(f"edge_attr has {t.size(1)} columns, layer expects {width}")
    return torch.cat([t, t.new_zeros(t.size(0), width - t.size(1))], dim=1)     # P7


class DynamicGraphLayer(nn.Module):
    """core/graph_layers.py:113-247. Returns node_dim features: LN(output_proj(gelu(conv2(gelu(conv1(x))))) + x)."""

    def __init__(self, node_dim, edge_dim, hidden_dim, num_heads=8, dropout=0.1, use_layer_norm=True):
        super().__init__()
        assert hidden_dim % num_heads == 0, "hidden_dim must be divisible by num_heads"
        self.node_dim, self.edge_dim, self.hidden_dim, self.num_heads = node_dim, edge_dim, hidden_dim, num_heads
        self.head_dim = hidden_dim // num_heads
        self.node_to_qkv = nn.Linear(node_dim, hidden_dim * 3)
        self.edge_to_key = nn.Linear(edge_dim, hidden_dim)
        self.graph_conv1 = GraphConvolution(node_dim, hidden_dim, edge_dim)
        self.graph_conv2 = GraphConvolution(hidden_dim, hidden_dim, edge_dim)
        self.output_proj = nn.Linear(hidden_dim, node_dim)
        self.dropout = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(node_dim) if use_layer_norm else nn.Identity()
        self.norm2 = nn.LayerNorm(node_dim) if use_layer_norm else nn.Identity()
        self.activation = nn.GELU()

    @torch.no_grad()
    def compute_dynamic_edges(self, x, edge_index, edge_attr, *, ctx=None):
        """Per-edge per-head attention weights, softmax over the in-edges of each target (dead inside forward, D14)."""
        _require_cuda(x)
        ctx = _ctx_for(x, edge_index, ctx)
        qkv = ops.linear(x, split_weight(self.node_to_qkv.weight), self.node_to_qkv.bias.detach())
        ek = ops.linear(_pad_cols(edge_attr, self.edge_dim), split_weight(self.edge_to_key.weight), self.edge_to_key.bias.detach())
        return ops.segment_softmax_scores(ctx.csr, qkv, ek, self.hidden_dim, self.num_heads)

    @torch.no_grad()
    def forward(self, x, edge_index=None, edge_attr=None, *, ctx=None, agg_attr=None, n_rows=None, post=None, part=None):
        """post: optional dict(transition=nn.Linear|None, ln=nn.LayerNorm|None, act=str) — the encoder tail
        (models/encoders.py:267-269 + P3) fused into the output_proj epilogue where shapes allow."""
        _require_cuda(x)
        ctx = _ctx_for(x, edge_index, ctx)
        if agg_attr is None and edge_attr is not None:
            agg_attr = ops.edge_attr_aggregate(ctx.csr if isinstance(ctx, GraphContext) else ctx, _pad_cols(edge_attr, self.edge_dim))
            if part is not None:
                agg_attr = part.local(agg_attr)
        h1 = self.graph_conv1(x, ctx=ctx, agg_attr=agg_attr, act="gelu", n_rows=n_rows, part=part)
        h2 = self.graph_conv2(h1, ctx=ctx, agg_attr=agg_attr, act="gelu", n_rows=n_rows, part=part)
        w = split_weight(self.output_proj.weight)
        b = self.output_proj.bias.detach()
        ln1 = _ln(self.norm1)
        if post is None:
            return ops.linear(h2, w, b, r_pre=x, ln1=ln1)
        tr, ln_out, act = post.get("transition"), post.get("ln"), post.get("act")
        if tr is None or isinstance(tr, nn.Identity):
            # LN(node_dim) directly followed by LN(out_dim) + act: one row-resident epilogue (identity (v))
            return ops.linear(h2, w, b, r_pre=x, ln1=ln1, ln2=_ln(ln_out), act_post=act)
        y = ops.linear(h2, w, b, r_pre=x, ln1=ln1)
        return ops.linear(y, split_weight(tr.weight), tr.bias.detach(), ln1=_ln(ln_out), act_post=act)


class AdaptiveGraphPooling(nn.Module):
    """core/graph_layers.py:250-329 (tanh scores, global top-k; `batch` is ignored exactly as in the reference, D16)."""

    def __init__(self, in_channels, ratio=0.5, min_score=None, multiplier=1.0, nonlinearity="tanh"):
        super().__init__()
        if min_score is not None or nonlinearity != "tanh" or multiplier != 1.0:
            raise NotImplementedError("only the reference defaults (ratio top-k, tanh, multiplier 1) are implemented")
        self.in_channels, self.ratio, self.min_score, self.multiplier = in_channels, ratio, min_score, multiplier
        self.score_net = nn.Sequential(nn.Linear(in_channels, in_channels // 2), nn.ReLU(), nn.Linear(in_channels // 2, 1))

    @torch.no_grad()
    def scores(self, x):
        l0, l2 = self.score_net[0], self.score_net[2]
        _, s = ops.linear(x, split_weight(l0.weight), l0.bias.detach(), act_post="relu", want_out=False,
                          dot=(l2.weight.detach().reshape(-1), float(_host_scalar(l2.bias)), "tanh"))
        return s

    @torch.no_grad()
    def pool(self, x, csr, act=None, part=None, with_eid=False):
        """Returns (pooled_x [k,C] with `act` applied, pooled CSR, perm int32 [k], node_map int32 [N], scores, pooled Part).
        Partitioned: scores are all-gathered, every rank runs the same exact top-k, and a selected node stays on its owner.
        with_eid: the pooled CSR keeps the level-0 edge ids (edge_attr[edge_mask] of :327 without the copy)."""
        s = self.scores(x)
        if part is None:
            k = max(1, int(self.ratio * x.size(0)))
            perm, node_map = ops.topk_select(s, k)
            px = ops.pool_gather(x, s, perm, act=act)
            return px, ops.csr_pool(csr, perm, node_map, with_eid), perm, node_map, s, None
        s_full = part.gather(s)
        k = max(1, int(self.ratio * part.n))
        perm, node_map = ops.topk_select(s_full, k)
        new_part = part.pooled(node_map)
        local_perm = (perm[new_part.a:new_part.b] - part.a).contiguous()
        px = ops.pool_gather(x, s, local_perm, act=act)
        return px, ops.csr_pool(csr, perm, node_map, with_eid), perm, node_map, s_full, new_part

    @torch.no_grad()
    def forward(self, x, edge_index, edge_attr=None, batch=None):
        """Reference signature: (pooled_x, pooled_edge_index, pooled_edge_attr, perm)."""
        _require_cuda(x)
        ctx = GraphContext(edge_index, None, x.size(0))
        px, _, perm, node_map, _, _ = self.pool(x, ctx.csr)
        nm = node_map.long()
        emask = (nm[edge_index[0]] >= 0) & (nm[edge_index[1]] >= 0)      # index bookkeeping for the public API only
        pei = nm[edge_index[:, emask]]
        pea = edge_attr[emask] if edge_attr is not None else None
        return px, pei, pea, perm.long()


_SCALAR_CACHE = _IdCache()


def _host_scalar(param):
    key = (param.data_ptr(), param._version)
    hit = _SCALAR_CACHE.get(param)
    if hit is None or hit[0] != key:
        hit = (key, float(param.detach().reshape(-1)[0].item()))
        _SCALAR_CACHE.put(param, hit)
    return hit[1]


class GraphUNet(nn.Module):
    """core/graph_layers.py:332-458 with P5 (edge_dim 32, up_convs on hidden_channels). D8 (decoder level j runs on
    the pooled edges of level j+1) is reproduced."""

    def __init__(self, in_channels, hidden_channels, out_channels, depth=3, pool_ratios=None, sum_res=True, act="relu"):
        super().__init__()
        if not sum_res:
            raise NotImplementedError("sum_res=False is not implemented (the reference default is True)")
        self.in_channels, self.hidden_channels, self.out_channels = in_channels, hidden_channels, out_channels
        self.depth, self.sum_res = depth, sum_res
        self.act_name = act if act in ("relu", "gelu") else "elu"
        pool_ratios = pool_ratios or [0.5] * depth
        mk = lambda nd: DynamicGraphLayer(node_dim=nd, edge_dim=32, hidden_dim=hidden_channels)
        self.down_convs = nn.ModuleList([mk(in_channels)] + [mk(hidden_channels) for _ in range(depth)])
        self.pools = nn.ModuleList([AdaptiveGraphPooling(hidden_channels, ratio=r) for r in pool_ratios])
        self.bottom_conv = mk(hidden_channels)
        self.up_convs = nn.ModuleList([mk(hidden_channels) for _ in range(depth)])
        self.final_conv = nn.Linear(hidden_channels, out_channels)

    @torch.no_grad()
    def forward(self, x, edge_index=None, edge_attr=None, batch=None, *, ctx=None, return_perms=False, part=None):
        """part (parallel.Part): x holds the owned rows of a node-partitioned graph; every level keeps its own Part."""
        _require_cuda(x)
        ctx = _ctx_for(x, edge_index, ctx)
        # edge attributes (:404-453 thread edge_attr through every level, pooled by `edge_attr[edge_mask]` :327): identity (i)
        # turns the per-edge term into one [N_l, 32] aggregate per level; the pooled CSRs keep the level-0 edge ids, so the
        # aggregate of level l reads the ORIGINAL edge_attr and nothing is compacted. No edges / no attributes -> no term
        # (edge_lin has no bias, P5 substitutes zeros).
        ea = _edge_attr_or_none(edge_attr, 32)
        csr0 = ctx.csr if isinstance(ctx, GraphContext) else ctx

        def agg_of(csr, lvl_part, n_lvl):
            """[rows owned at this level that exist in `csr`, 32] aggregate (D8: the decoder's CSR has fewer rows than nodes)."""
            if ea is None:
                return None
            g = _level0_agg(ctx, ea) if csr is csr0 else ops.edge_attr_aggregate(csr, ea)
            if lvl_part is None:
                return g
            return g[min(lvl_part.a, csr.N):min(lvl_part.b, csr.N)]

        a = self.act_name
        if part is not None:
            x = self._forward_partitioned(x, ctx, csr0, ea, part, a)
            out = ops.linear(x, split_weight(self.final_conv.weight), self.final_conv.bias.detach())
            return (out, self._last_perms) if return_perms else out
        csrs, xs, maps, perms = [csr0], [], [], []
        aggs = [agg_of(csr0, None, csr0.N)]
        x = self.down_convs[0](x, ctx=csrs[0], agg_attr=aggs[0])
        xs.append(x)
        # `cur` is act(x) of the reference (:419/:432) — for level 0 applied by a tiny un-pool kernel with an empty map
        cur = ops.unpool_add_act(x, x, _neg_map(x), act=a)
        for i in range(self.depth):
            x = self.down_convs[i + 1](cur, ctx=csrs[-1], agg_attr=aggs[-1])
            xs.append(x)
            cur, pcsr, perm, node_map, _, _ = self.pools[i].pool(x, csrs[-1], act=a, with_eid=ea is not None)   # pooled + activated
            csrs.append(pcsr); maps.append(node_map); perms.append(perm)
            aggs.append(agg_of(pcsr, None, pcsr.N))
        x = self.bottom_conv(cur, ctx=csrs[-1], agg_attr=aggs[-1])
        for i in range(self.depth):
            j = self.depth - 1 - i
            cur = ops.unpool_add_act(xs[j + 1], x, maps[j], act=a)
            # D8: level-(j+1) edges and attributes on the level-j node set; rows beyond the pooled CSR have no in-edges
            x = self.up_convs[i](cur, ctx=csrs[j + 1], agg_attr=aggs[j + 1], n_rows=cur.size(0))
        out = ops.linear(x, split_weight(self.final_conv.weight), self.final_conv.bias.detach())
        return (out, perms) if return_perms else out

    def _forward_partitioned(self, x, ctx, csr0, ea, part, a):
        return _unet_forward_partitioned(self, x, ctx, csr0, ea, part, a)


def _unet_forward_partitioned(self, x, ctx, csr0, ea, part, a):
    """Node-partitioned U-Net. Only LEVEL 0 (the full node set: down_convs[0], down_convs[1] and the last up conv — 6 of the 16
    propagates, > 80 % of the U-Net's bytes) is row-partitioned with a halo exchange per propagate. The pooled levels (1/2, 1/4,
    1/8 of the nodes, 128-wide) are REPLICATED: every rank all-gathers the level-0 output once, runs the same exact top-k and the
    small levels redundantly — 10 propagates on <= N/2 rows cost less than the 10 latency-bound exchanges (+ 3 host reads for the
    uneven pooled bounds) they replace, and their results are bitwise those of a single GPU."""
    def agg_level0():
        if ea is None:
            return None
        return part.local(_level0_agg(ctx, ea))
    agg0 = agg_level0()
    x0 = self.down_convs[0](x, ctx=csr0, agg_attr=agg0, part=part)
    cur = ops.unpool_add_act(x0, x0, _neg_map(x0), act=a)
    x1 = self.down_convs[1](cur, ctx=csr0, agg_attr=agg0, part=part)            # skip tensor of level 0 (owned rows)
    x1_full = part.gather(x1)
    # ---- replicated from here: pool level 0 -> 1 and everything below, exactly the single-GPU code on full tensors
    csrs, xs, maps, perms, aggs = [csr0], [None, x1_full], [], [], [None]
    xl = x1_full
    for i in range(self.depth):
        if i > 0:
            xl = self.down_convs[i + 1](cur, ctx=csrs[-1], agg_attr=aggs[-1])
            xs.append(xl)
        cur, pcsr, perm, node_map, _, _ = self.pools[i].pool(xl, csrs[-1], act=a, with_eid=ea is not None)
        csrs.append(pcsr); maps.append(node_map); perms.append(perm)
        aggs.append(None if ea is None else ops.edge_attr_aggregate(pcsr, ea))
    xl = self.bottom_conv(cur, ctx=csrs[-1], agg_attr=aggs[-1])
    for i in range(self.depth - 1):
        j = self.depth - 1 - i
        cur = ops.unpool_add_act(xs[j + 1], xl, maps[j], act=a)
        xl = self.up_convs[i](cur, ctx=csrs[j + 1], agg_attr=aggs[j + 1], n_rows=cur.size(0))
    # ---- back to the partition for the level-0 decoder (D8: level-1 edges on the level-0 node set)
    nm = part.local(maps[0]).contiguous()
    cur = ops.unpool_add_act(x1, xl, nm, act=a)
    agg_d = None if ea is None else aggs[1][min(part.a, csrs[1].N):min(part.b, csrs[1].N)]
    self._last_perms = perms
    return self.up_convs[self.depth - 1](cur, ctx=csrs[1], agg_attr=agg_d, n_rows=cur.size(0), part=part)


def _edge_attr_or_none(edge_attr, width):
    """fp32, zero-padded to `width` columns (P7); None when there are no attributes or no edges."""
    if edge_attr is None or edge_attr.size(0) == 0:
        return None
    return _pad_cols(edge_attr.float(), width)


def _level0_agg(ctx, ea):
    """sum_e norm_e * edge_attr_e per destination of the input graph, shared by the encoder and the U-Net (cached on the
    GraphContext, keyed on the attribute tensor)."""
    if not isinstance(ctx, GraphContext):
        return ops.edge_attr_aggregate(ctx, ea)
    key = (ea.data_ptr(), ea._version, tuple(ea.shape))
    hit = ctx.agg_attr
    if hit is None or hit[0] != key:
        hit = (key, ops.edge_attr_aggregate(ctx.csr, ea), ea)      # `ea` kept alive: the key holds its address
        ctx.agg_attr = hit
    return hit[1]


_NEG_MAPS = {}


def _neg_map(x):
    key = (x.device, x.size(0))
    m = _NEG_MAPS.get(key)
    if m is None:
        if len(_NEG_MAPS) > 8:
            _NEG_MAPS.clear()
        m = torch.full((x.size(0),), -1, dtype=torch.int32, device=x.device)
        _NEG_MAPS[key] = m
    return m


# ------------------------------------------------------------------------------------------ attention
class MultiHeadAttention(nn.Module):
    """core/attention.py:16-181. Projections on tcgen05; the core runs flash-style when no weights are requested
    and no arbitrary mask is given, otherwise through the dense-weights kernel (O(L^2) output anyway)."""

    def __init__(self, embed_dim, num_heads, dropout=0.1, bias=True, kdv_bias=True, batch_first=True, add_zero_attn=False):
        super().__init__()
        if add_zero_attn:
            raise NotImplementedError("add_zero_attn=True is not implemented (reference default False)")
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        self.batch_first, self.add_zero_attn = batch_first, add_zero_attn
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=kdv_bias)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=kdv_bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.attn_dropout, self.resid_dropout = nn.Dropout(dropout), nn.Dropout(dropout)
        for l in (self.q_proj, self.k_proj, self.v_proj, self.out_proj):
            nn.init.xavier_uniform_(l.weight)
            if l.bias is not None:
                nn.init.constant_(l.bias, 0.0)

    def qkv_weight(self):
        lin = [self.q_proj, self.k_proj, self.v_proj]
        return split_weight_cat([l.weight for l in lin], "qkv"), _qkv_bias(lin)

    def kv_weight(self):
        lin = [self.k_proj, self.v_proj]
        return split_weight_cat([l.weight for l in lin], "kv"), _qkv_bias(lin)

    def _project(self, query, key, value):
        """q / k / v projections as views of fused GEMM outputs. Returns (q [B*Lq, D], k [B*Lk, D], v [B*Lk, D], qkv | None,
        kv_qkv | None) — the last two are the [*, 3D] q|k|v-layout buffers the flash kernels read (kv_qkv None = self-attention)."""
        D = self.embed_dim
        q2 = query.reshape(-1, D).contiguous()
        if key is None or key is query:
            w, bias = self.qkv_weight()
            qkv = ops.linear(q2, w, bias)
            return qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], qkv, None
        k2 = key.reshape(-1, D).contiguous()
        bq = self.q_proj.bias.detach() if self.q_proj.bias is not None else None
        qkv = torch.empty(q2.size(0), 3 * D, dtype=torch.float32, device=q2.device)      # only the q third is written / read
        ops.linear(q2, split_weight(self.q_proj.weight), bq, out=qkv[:, :D])
        kv = torch.empty(k2.size(0), 3 * D, dtype=torch.float32, device=q2.device)
        if value is None or value is key:
            w, bias = self.kv_weight()
            ops.linear(k2, w, bias, out=kv[:, D:])
        else:
            v2 = value.reshape(-1, D).contiguous()
            if v2.size(0) != k2.size(0):
                raise ops.KernelError(f"key has {k2.size(0)} rows but value has {v2.size(0)}")
            ops.linear(k2, split_weight(self.k_proj.weight), self.k_proj.bias.detach() if self.k_proj.bias is not None else None,
                       out=kv[:, D:2 * D])
            ops.linear(v2, split_weight(self.v_proj.weight), self.v_proj.bias.detach() if self.v_proj.bias is not None else None,
                       out=kv[:, 2 * D:])
        return qkv[:, :D], kv[:, D:2 * D], kv[:, 2 * D:], qkv, kv

    @torch.no_grad()
    def forward(self, query, key=None, value=None, key_padding_mask=None, need_weights=True, attn_mask=None,
                average_attn_weights=True):
        """core/attention.py:73-181 (eval: both dropouts are identities). Without weights and masks the core is the flash
        kernel (all batch elements in one launch, one segment each); weights or masks go through the dense kernel
        (dgdm_mha_dense), whose output is O(L_q L_k) anyway."""
        _require_cuda(query)
        if not self.batch_first:
            query = query.transpose(0, 1)
            key = key.transpose(0, 1) if key is not None else None
            value = value.transpose(0, 1) if value is not None else None
        B, L, D = query.shape
        o, w = self._core(query, key, value, key_padding_mask, need_weights, attn_mask, average_attn_weights)
        out = ops.linear(o, split_weight(self.out_proj.weight), self.out_proj.bias.detach() if self.out_proj.bias is not None else None)
        out = out.view(B, L, D)
        if not self.batch_first:
            out = out.transpose(0, 1)
        return out, w

    def _core(self, query, key, value, key_padding_mask, need_weights, attn_mask, average_attn_weights, *, _pos=None,
              _pos_by_head=False, _temperature=1.0):
        """Projections + softmax(QK^T/sqrt(hd) + masks) V on batch-first inputs; returns (o [B*L, D] before out_proj, weights).
        `_pos*`: SpatialAttention's distance bias, formed inside the kernel instead of as a [B, L, L] tensor."""
        B, L, D = query.shape
        Lk = L if key is None else key.size(1)
        q, k, v, qkv, kv = self._project(query, key, value)
        dense = need_weights or attn_mask is not None or key_padding_mask is not None or _pos is not None \
            or self.head_dim not in (4, 8, 16, 32, 64)
        w = None
        if dense:
            o, w = ops.mha_dense(q, k, v, B, L, Lk, self.num_heads, attn_mask=attn_mask, key_padding_mask=key_padding_mask,
                                 pos_q=_pos, pos_k=_pos, pos_by_head=_pos_by_head, temperature=_temperature,
                                 need_weights=need_weights, per_head=not average_attn_weights)
        else:
            dev = query.device
            qseg = torch.arange(0, (B + 1) * L, L, dtype=torch.int32, device=dev)
            kseg = torch.arange(0, (B + 1) * Lk, Lk, dtype=torch.int32, device=dev)
            o = ops.spatial_attention_any(qkv, None, qseg, B, self.num_heads, kv_qkv=kv, kseg_ptr=kseg, max_seg=Lk)
        return o, w


def _qkv_bias(linears):
    if all(l.bias is None for l in linears):
        return None
    parts = [l.bias.detach() if l.bias is not None else l.weight.new_zeros(l.weight.size(0)) for l in linears]
    return torch.cat(parts)


class SpatialAttention(nn.Module):
    """core/attention.py:184-327. LN(x + MHA(x + PE(pos), bias = -|p_i - p_j| / T)); the N x N bias and scores are never formed."""

    # return_attention materialises [n, n] head-averaged weights: refuse above this many nodes per graph
    MAX_WEIGHT_NODES = 16384

    def __init__(self, embed_dim, num_heads=8, max_positions=10000, dropout=0.1, temperature=1.0):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.max_positions, self.temperature = max_positions, temperature
        self.attention = MultiHeadAttention(embed_dim, num_heads, dropout=dropout)
        self.pos_encoding = nn.Parameter(torch.randn(max_positions, embed_dim) * 0.02)       # dead in the reference (D15); kept for state_dict
        self.spatial_proj = nn.Sequential(nn.Linear(2, embed_dim // 2), nn.ReLU(), nn.Linear(embed_dim // 2, embed_dim))
        self.norm = nn.LayerNorm(embed_dim)

    @torch.no_grad()
    def forward_nodes(self, x, pos, seg_ptr, num_seg, seg_host=None, need_weights=False, *, part=None):
        """x [N,D], pos [N,2] for a node-sorted batch with segment pointers. Returns (out [N,D], weights list|None).
        part (parallel.Part): x holds the owned rows, pos / seg_ptr describe the whole batch; queries = owned rows,
        keys / values = the all-gathered q|k|v rows (context parallelism)."""
        att = self.attention
        if self.head_dim not in (4, 8, 16, 32, 64):
            raise NotImplementedError(f"SpatialAttention head_dim {self.head_dim}: supported head widths are 4, 8, 16, 32, 64")
        pos = pos.float().contiguous()
        w, bias = att.qkv_weight()
        max_seg = max((seg_host[i + 1] - seg_host[i] for i in range(num_seg)), default=0) if seg_host is not None else 0
        if part is None:
            xp = ops.spatial_pe_add(x, pos, seg_ptr, num_seg)
            qkv = ops.linear(xp, w, bias)
            o = ops.spatial_attention_any(qkv, pos, seg_ptr, num_seg, self.num_heads, self.temperature, max_seg=max_seg)
        else:
            if need_weights:
                raise NotImplementedError("return_attention is not available for node-partitioned graphs")
            pos_l = part.local(pos).contiguous()
            qseg = (torch.clamp(seg_ptr, part.a, part.b) - part.a).to(torch.int32)
            xp = ops.spatial_pe_add(x, pos_l, qseg, num_seg, minmax_from=(pos, seg_ptr))
            qkv = ops.linear(xp, w, bias)
            qkv_full = part.gather(qkv)
            o = ops.spatial_attention_any(qkv, pos_l, qseg, num_seg, self.num_heads, self.temperature,
                                          kv_qkv=qkv_full, kv_pos=pos, kseg_ptr=seg_ptr, max_seg=max_seg)
        out = ops.linear(o, split_weight(att.out_proj.weight), att.out_proj.bias.detach(), r_pre=x, ln1=_ln(self.norm))
        weights = None
        if need_weights:
            weights = []
            for g in range(num_seg):
                a, b = seg_host[g], seg_host[g + 1]
                if b - a > self.MAX_WEIGHT_NODES:
                    raise ops.KernelError(f"attention weights for a {b - a}-node graph would need {(b - a) ** 2 * 4 / 1e9:.1f} GB; "
                                          f"return_attention is limited to {self.MAX_WEIGHT_NODES} nodes per graph")
                weights.append(ops.spatial_attention_weights(qkv[a:b], pos[a:b], self.num_heads, self.temperature))
        return out, weights

    @torch.no_grad()
    def forward(self, x, positions, mask=None):
        """Reference signature (core/attention.py:285-327): x [B,n,D], positions [B,n,2] -> (out [B,n,D], weights [B,n,n]).
        The PE min/max is taken over the WHOLE positions tensor (:233-236), not per batch element. The bias is added as
        `[B,H,n,n] += bias[B,n,n]`, which torch only broadcasts for B == 1 — or, by accident, B == H, where element h of the
        bias lands on HEAD h of every batch element; both behaviours are reproduced, any other B raises like the reference."""
        _require_cuda(x)
        B, n, D = x.shape
        if B != 1 and B != self.num_heads:
            raise RuntimeError(f"The size of tensor a ({self.num_heads}) must match the size of tensor b ({B}) at non-singleton dimension 1")
        x2 = x.reshape(B * n, D).contiguous()
        pos = positions.reshape(B * n, 2).float().contiguous()
        one = torch.tensor([0, B * n], dtype=torch.int32, device=x.device)
        xp = ops.spatial_pe_add(x2, pos, one, 1)                                    # global min / max: a single segment
        att = self.attention
        o, w = att._core(xp.view(B, n, D), None, None, None, True, mask, True, _pos=pos.view(B, n, 2), _pos_by_head=B != 1,
                         _temperature=self.temperature)
        out = ops.linear(o, split_weight(att.out_proj.weight), att.out_proj.bias.detach(), r_pre=x2, ln1=_ln(self.norm))
        return out.view(B, n, D), w


# ------------------------------------------------------------------------------------------ diffusion
class DiffusionScheduler:
    """core/diffusion.py:16-61 — host-side tables (fp32, same torch ops as the reference so the values are identical)."""

    def __init__(self, num_timesteps=1000, beta_start=0.0001, beta_end=0.02, schedule="cosine"):
        self.num_timesteps, self.schedule = num_timesteps, schedule
        if schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_timesteps)
        elif schedule == "cosine":
            self.betas = self._cosine_beta_schedule(num_timesteps)
        elif schedule == "sigmoid":
            self.betas = self._sigmoid_beta_schedule(num_timesteps, beta_start, beta_end)
        else:
            raise ValueError(f"Unknown schedule: {schedule}")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alphas_cumprod_prev = F.pad(self.alphas_cumprod[:-1], (1, 0), value=1.0)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)

    def _cosine_beta_schedule(self, timesteps, s=0.008):
        x = torch.linspace(0, timesteps, timesteps + 1)
        ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)

    def _sigmoid_beta_schedule(self, timesteps, start=-3, end=3):
        return torch.sigmoid(torch.linspace(-6, 6, timesteps)) * (end - start) + start


class DiffusionLayer(nn.Module):
    """core/diffusion.py:64-275. 2-D [N, D] inputs (P6: the reference's GroupNorm rejects 3-D); [1,N,D] is squeezed."""

    def __init__(self, node_dim, hidden_dim, num_timesteps=1000, schedule="cosine", conditioning_dim=None):
        super().__init__()
        self.node_dim, self.hidden_dim, self.num_timesteps = node_dim, hidden_dim, num_timesteps
        self.scheduler = DiffusionScheduler(num_timesteps, schedule=schedule)
        self.time_embed = nn.Sequential(nn.Linear(128, hidden_dim), nn.SiLU(), nn.Linear(hidden_dim, hidden_dim))
        self.denoise_net = nn.Sequential(
            nn.Linear(node_dim + hidden_dim, hidden_dim * 2), nn.GroupNorm(8, hidden_dim * 2), nn.SiLU(), nn.Dropout(0.1),
            nn.Linear(hidden_dim * 2, hidden_dim), nn.GroupNorm(8, hidden_dim), nn.SiLU(), nn.Dropout(0.1),
            nn.Linear(hidden_dim, node_dim))
        self.condition_net = nn.Linear(conditioning_dim, hidden_dim) if conditioning_dim is not None else None
        self._tables = {}
        self._graphs = {}

    def get_timestep_embedding(self, timesteps, dim=128):
        half = dim // 2
        e = math.log(10000) / (half - 1)
        e = torch.exp(torch.arange(half, device=timesteps.device) * -e)
        e = timesteps.float()[:, None] * e[None, :]
        e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
        if dim % 2 == 1:
            e = torch.cat([e, torch.zeros_like(e[:, :1])], dim=1)
        return e

    def _dev_tables(self, device):
        t = self._tables.get(device)
        if t is None:
            ac = self.scheduler.alphas_cumprod
            t = (torch.sqrt(ac).to(device), torch.sqrt(1.0 - ac).to(device))
            self._tables[device] = t
        return t

    @staticmethod
    def _segments(timesteps, n, device):
        """timesteps [B] (one per graph segment given by equal split) or [N] (per node, constant per run)."""
        t = timesteps.reshape(-1)
        if t.numel() == n and n > 1:
            th = t.cpu()
            change = torch.nonzero(th[1:] != th[:-1]).reshape(-1) + 1
            starts = [0] + change.tolist()
            seg = starts + [n]
            seg_t = th[starts]
        else:
            if t.numel() != 1:
                raise ops.KernelError("timesteps must have one entry per node or a single entry")
            seg, seg_t = [0, n], t.cpu()
        return (torch.tensor(seg, dtype=torch.int32, device=device), seg_t.to(torch.int32).to(device), seg_t.long())

    @torch.no_grad()
    def add_noise(self, x_start, noise, timesteps):
        _require_cuda(x_start)
        x2 = x_start.reshape(-1, x_start.size(-1)).contiguous()
        seg_ptr, seg_t, _ = self._segments(timesteps, x2.size(0), x2.device)
        sa, sb = self._dev_tables(x2.device)
        return ops.diffusion_add_noise(x2, noise.reshape(x2.shape).contiguous(), sa, sb, seg_ptr, seg_t).view(x_start.shape)

    @torch.no_grad()
    def predict_noise(self, x_noisy, timesteps, condition=None):
        """concat([x_t, t_emb]) W^T = x_t W_x^T + t_emb W_t^T with t_emb constant per segment -> per-segment bias row
        (identity (iv)); GroupNorm(8)+SiLU row-wise; three tcgen05 GEMMs."""
        _require_cuda(x_noisy)
        shp = x_noisy.shape
        x2 = x_noisy.reshape(-1, shp[-1]).contiguous()
        n = x2.size(0)
        seg_ptr, seg_t32, seg_t = self._segments(timesteps, n, x2.device)
        te0, te2 = self.time_embed[0], self.time_embed[2]
        d0, g1, d4, g5, d8 = (self.denoise_net[i] for i in (0, 1, 4, 5, 8))
        emb = self.get_timestep_embedding(seg_t.to(x2.device))                                  # [S,128] tiny, host-driven
        nseg = emb.size(0)
        t1 = ops.linear(emb.contiguous(), split_weight(te0.weight), te0.bias.detach(), act_post="silu")
        t2 = ops.linear(t1, split_weight(te2.weight), te2.bias.detach())                        # [S, hidden]
        wx, wt = _split_cols(d0.weight, self.node_dim)
        tb = ops.linear(t2, wt, d0.bias.detach())                                               # [S, 2*hidden] per-segment bias rows
        seg_h = seg_ptr.cpu().tolist()
        row_idx = torch.repeat_interleave(torch.arange(nseg, device=x2.device, dtype=torch.int32),
                                          torch.tensor([seg_h[i + 1] - seg_h[i] for i in range(nseg)], device=x2.device))
        cond_rows = None
        if condition is not None and self.condition_net is not None:
            # t_emb + condition_net(condition) (core/diffusion.py:158-160): the extra term goes through the time-embedding columns of the
            # first denoise GEMM as well -> one [N, cdim] x [cdim, hidden] and one [N, hidden] x [hidden, 2 hidden] GEMM, added as a residual
            c2 = condition.reshape(-1, condition.size(-1)).float()
            if c2.size(0) == 1 and n > 1:
                c2 = c2.expand(n, -1)
            if c2.size(0) != n:
                raise ops.KernelError(f"condition has {c2.size(0)} rows for {n} nodes")
            ce = ops.linear(c2.contiguous(), split_weight(self.condition_net.weight), self.condition_net.bias.detach())
            cond_rows = ops.linear(ce, wt, None)
        h = ops.linear(x2, wx, None, row_bias=(row_idx.contiguous(), tb), r_pre=cond_rows)
        h = ops.groupnorm_silu(h, 8, g1.weight.detach(), g1.bias.detach(), g1.eps)
        h = ops.linear(h, split_weight(d4.weight), d4.bias.detach())
        h = ops.groupnorm_silu(h, 8, g5.weight.detach(), g5.bias.detach(), g5.eps)
        out = ops.linear(h, split_weight(d8.weight), d8.bias.detach())
        return out.view(shp)

    @torch.no_grad()
    def forward(self, x_start, timesteps=None, noise=None, condition=None):
        _require_cuda(x_start)
        squeeze = x_start.dim() == 3 and x_start.size(0) == 1
        x2 = x_start[0] if squeeze else x_start
        if x2.dim() != 2:
            raise ops.KernelError("DiffusionLayer runs on [N, D] (or [1, N, D]) inputs (reference GroupNorm limitation, D9/P6)")
        if timesteps is None:
            timesteps = torch.randint(0, self.num_timesteps, (1,), device=x2.device)
        if noise is None:
            noise = torch.randn_like(x2)
        noise = noise.reshape(x2.shape)
        t = timesteps.reshape(-1)
        if t.numel() not in (1, x2.size(0)):
            t = t[:1]
        xn = self.add_noise(x2, noise, t)
        pn = self.predict_noise(xn, t, condition)
        return (xn.unsqueeze(0), pn.unsqueeze(0)) if squeeze else (xn, pn)

    def _step_bias_rows(self, ts, device):
        """Per-step bias rows of the first denoise GEMM for ALL sampling steps at once (identity (iv): concat([x_t, t_emb]) W^T =
        x_t W_x^T + t_emb W_t^T with t_emb constant per step): three tiny GEMMs instead of three per step."""
        te0, te2, d0 = self.time_embed[0], self.time_embed[2], self.denoise_net[0]
        emb = self.get_timestep_embedding(ts.to(device))                                            # [S,128]
        t1 = ops.linear(emb.contiguous(), split_weight(te0.weight), te0.bias.detach(), act_post="silu")
        t2 = ops.linear(t1, split_weight(te2.weight), te2.bias.detach())
        _, wt = _split_cols(d0.weight, self.node_dim)
        return ops.linear(t2, wt, d0.bias.detach())                                                 # [S, 2*hidden]

    def _sample_loop(self, x, tb_all, row_idx, ts_host, noises):
        """The DDPM loop on device tensors only (no host reads): capturable into one CUDA graph."""
        d0, g1, d4, g5, d8 = (self.denoise_net[i] for i in (0, 1, 4, 5, 8))
        wx, _ = _split_cols(d0.weight, self.node_dim)
        sch = self.scheduler
        last = len(ts_host) - 1
        for i, ti in enumerate(ts_host):
            h = ops.linear(x, wx, None, row_bias=(row_idx, tb_all[i:i + 1]))
            h = ops.groupnorm_silu(h, 8, g1.weight.detach(), g1.bias.detach(), g1.eps)
            h = ops.linear(h, split_weight(d4.weight), d4.bias.detach())
            h = ops.groupnorm_silu(h, 8, g5.weight.detach(), g5.bias.detach(), g5.eps)
            eps = ops.linear(h, split_weight(d8.weight), d8.bias.detach())
            x = ops.ddpm_update(x, eps, None if i == last else noises[i], float(torch.sqrt(1 - sch.alphas_cumprod[ti])),
                                float(1.0 / torch.sqrt(sch.alphas_cumprod[ti])), float(torch.sqrt(sch.alphas[ti])),
                                float(torch.sqrt(sch.posterior_variance[ti])))
        return x

    @torch.no_grad()
    def sample(self, shape, device, condition=None, num_inference_steps=50, *, init=None, step_noise=None, use_graph=True):
        """DDPM loop (core/diffusion.py:214-275). init / step_noise inject the random draws for parity tests.
        The reference issues ~25 tiny torch ops per step (launch-bound); here a step is 6 kernels, the per-step time-embedding
        GEMMs are hoisted out of the loop, and the whole loop is captured once per (shape, steps) into a CUDA graph and replayed."""
        if condition is not None:
            raise NotImplementedError("conditioning is not used by DGDMModel and is not implemented")
        device = torch.device(device)
        shape = tuple(shape)
        x0 = torch.randn(shape, device=device) if init is None else init.to(device).float()
        x0 = x0.reshape(-1, shape[-1]).contiguous()
        n = x0.size(0)
        ts = torch.linspace(self.num_timesteps - 1, 0, num_inference_steps, dtype=torch.long)
        ts_host = [int(t) for t in ts]
        steps = len(ts_host)
        if step_noise is None:
            noise = torch.randn(max(steps - 1, 1), n, shape[-1], device=device)
        else:
            noise = torch.stack([step_noise[i].to(device).float().reshape(n, shape[-1]) for i in range(steps - 1)]) if steps > 1 \
                else torch.zeros(1, n, shape[-1], device=device)
        key = (n, shape[-1], steps, device.index)
        entry = self._graphs.get(key) if use_graph else None
        if entry is None:
            tb_all = self._step_bias_rows(ts, device)
            row_idx = torch.zeros(n, dtype=torch.int32, device=device)
            if not use_graph:
                return self._sample_loop(x0, tb_all, row_idx, ts_host, noise).view(shape)
            sx, sn = x0.clone(), noise.clone()
            self._sample_loop(sx, tb_all, row_idx, ts_host, sn)                    # warm-up: weight planes, kernel attributes
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._sample_loop(sx, tb_all, row_idx, ts_host, sn)
            entry = (graph, sx, sn, out, tb_all, row_idx, self._weights_version())
            self._graphs[key] = entry
        graph, sx, sn, out, _, _, ver = entry
        if ver != self._weights_version():          # parameters changed since capture: drop the cache and start over
            self._graphs.clear()
            return self.sample(shape, device, None, num_inference_steps, init=x0.view(shape), step_noise=[noise[i] for i in range(steps - 1)],
                               use_graph=use_graph)
        sx.copy_(x0)
        sn.copy_(noise)
        graph.replay()
        return out.clone().view(shape)

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())


_COLSPLIT_CACHE = _IdCache()


def _split_cols(param, k0):
    """SplitWeights of W[:, :k0] and W[:, k0:] (denoise_net.0 acts on concat([x_t, t_emb]))."""
    key = (param.data_ptr(), param._version, k0)
    hit = _COLSPLIT_CACHE.get(param)
    if hit is None or hit[0] != key:
        w = param.detach()
        hit = (key, ops.SplitWeight(w[:, :k0].contiguous()), ops.SplitWeight(w[:, k0:].contiguous()))
        _COLSPLIT_CACHE.put(param, hit)
    return hit[1], hit[2]


# ------------------------------------------------------------------------------------------ encoders
_ACTS = {"relu": nn.ReLU, "gelu": nn.GELU, "elu": nn.ELU}


class FeatureEncoder(nn.Module):
    """models/encoders.py:19-124. (Linear -> LayerNorm -> act) x num_layers + residual projection; normalization 'layer'
    or 'none' (batch / instance norm are not benchmarked and raise)."""

    def __init__(self, input_dim, hidden_dim, num_layers=2, dropout=0.1, activation="gelu", normalization="layer", use_residual=True):
        super().__init__()
        if activation not in _ACTS:
            raise ValueError(f"Unknown activation: {activation}")
        self.input_dim, self.hidden_dim, self.num_layers, self.use_residual = input_dim, hidden_dim, num_layers, use_residual
        self.activation_name = activation
        self.activation = _ACTS[activation]()
        mk_norm = {"layer": lambda: nn.LayerNorm(hidden_dim), "batch": lambda: nn.BatchNorm1d(hidden_dim),
                   "instance": lambda: nn.InstanceNorm1d(hidden_dim)}.get(normalization, lambda: nn.Identity())      # encoders.py:93-102
        layers = [nn.Linear(input_dim, hidden_dim), mk_norm(), self.activation, nn.Dropout(dropout)]
        for _ in range(num_layers - 1):
            layers += [nn.Linear(hidden_dim, hidden_dim), mk_norm(), self.activation, nn.Dropout(dropout)]
        self.encoder = nn.Sequential(*layers)
        self.residual_proj = nn.Linear(input_dim, hidden_dim) if use_residual and input_dim != hidden_dim else None

    @torch.no_grad()
    def forward(self, x):
        _require_cuda(x)
        x = x.contiguous()
        res = None
        if self.use_residual:
            res = x if self.residual_proj is None else ops.linear(x, split_weight(self.residual_proj.weight), self.residual_proj.bias.detach())
        h = x
        for i in range(self.num_layers):
            lin, norm = self.encoder[4 * i], self.encoder[4 * i + 1]
            last = i == self.num_layers - 1
            if isinstance(norm, nn.BatchNorm1d):
                h = _norm_act_torch(ops.linear(h, split_weight(lin.weight), lin.bias.detach()), norm, self.activation_name, self.training)
                if last and res is not None:
                    h = h + res
                continue
            h = ops.linear(h, split_weight(lin.weight), lin.bias.detach(), ln1=_ln(norm), act_post=self.activation_name,
                           r_post=res if last else None)
        return h


class GraphEncoder(nn.Module):
    """models/encoders.py:127-280 with P1 (index clamp) and P3 (transitions)."""

    def __init__(self, input_dim, hidden_dims, num_layers=4, attention_heads=8, dropout=0.1, activation="gelu",
                 normalization="layer", use_edge_features=True, aggregation="add"):
        super().__init__()
        if not use_edge_features:
            raise NotImplementedError("use_edge_features=False is not implemented (reference default True)")
        self.input_dim, self.hidden_dims, self.num_layers, self.use_edge_features = input_dim, hidden_dims, num_layers, use_edge_features
        dims = [input_dim] + list(hidden_dims)
        self.graph_layers, self.norm_layers, trans = nn.ModuleList(), nn.ModuleList(), []
        for i in range(num_layers):
            a, b = dims[min(i, len(dims) - 1)], dims[min(i + 1, len(dims) - 1)]
            self.graph_layers.append(DynamicGraphLayer(node_dim=a, edge_dim=32, hidden_dim=b, num_heads=attention_heads, dropout=dropout))
            self.norm_layers.append({"layer": nn.LayerNorm, "batch": nn.BatchNorm1d, "instance": nn.InstanceNorm1d}[normalization](b)
                                    if normalization in ("layer", "batch", "instance") else nn.Identity())            # encoders.py:217-226
            trans.append(nn.Linear(a, b) if a != b else nn.Identity())
        self.transitions = nn.ModuleList(trans)
        self.activation_name = activation if activation in _ACTS else "relu"
        self.activation = _ACTS[self.activation_name]()
        self.dropout = nn.Dropout(dropout)
        self.output_proj = nn.Linear(dims[-1], dims[-1])

    @torch.no_grad()
    def forward(self, x, edge_index=None, edge_attr=None, batch=None, *, ctx=None, part=None):
        _require_cuda(x)
        ctx = _ctx_for(x, edge_index, ctx)
        agg = None
        ea = _edge_attr_or_none(edge_attr, 32) if self.use_edge_features else None
        if ea is not None:
            agg = _level0_agg(ctx, ea)
            if part is not None:
                agg = part.local(agg)
        h, outs = x, []
        for i, layer in enumerate(self.graph_layers):
            norm = self.norm_layers[i]
            if isinstance(norm, nn.BatchNorm1d):
                h = layer(h, ctx=ctx, agg_attr=agg, part=part, post=dict(transition=self.transitions[i], ln=None, act=None))
                h = _norm_act_torch(h, norm, self.activation_name, self.training)
            else:
                h = layer(h, ctx=ctx, agg_attr=agg, part=part,
                          post=dict(transition=self.transitions[i], ln=norm, act=self.activation_name))
            outs.append(h)
        final = ops.linear(h, split_weight(self.output_proj.weight), self.output_proj.bias.detach())
        return {"embeddings": final, "layer_outputs": outs, "num_nodes": x.size(0)}
