"""Differentiable forward of DGDMModel for the training entry points (`pretrain_step`, mode='pretrain' /
'finetune' under autograd; reference: models/dgdm_model.py:405-506, training/trainer.py:100-128).

The inference forward fuses every row-wise epilogue into its GEMM and keeps no intermediates; a backward pass needs them, so
training runs this restatement of the same layers as an autograd graph whose heavy nodes are hand-written kernels in BOTH
directions:
  * `linear`            forward dgdm_gemm_f16x2 (split-fp16 tcgen05); backward dX through the same kernel on W^T,
                        dW = dY^T X as a plain library GEMM (cuBLAS via torch.matmul), db a column sum;
  * `gcn_propagate`     forward dgdm_spmm_gcn; backward the same kernel on the TRANSPOSED CSR (by source), fixed order;
  * `spatial_attention` forward the tcgen05 flash kernel; backward dgdm_spatial_attention_backward (three recompute passes).
LayerNorm / GroupNorm / activations / residual adds / gathers are element- and row-wise torch autograd ops in this path
(their fused inference versions have no backward yet). Dropout follows `model.training` where the layer exposes it; the
attention-weight dropout inside the flash kernel is not applied (train with the reference's eval semantics for that one op).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .graph import GraphContext, get_context

_ACT = {"gelu": F.gelu, "relu": F.relu, "elu": F.elu, "silu": F.silu}


def _gemm_ok(n_out, k):
    return n_out >= 16 and n_out % 16 == 0 and k >= 4 and k % 4 == 0


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        if not _gemm_ok(weight.size(0), weight.size(1)):              # a few tiny layers (e.g. the 64 -> 1 pooling score)
            return F.linear(x, weight, bias)
        return ops.linear(x, ops.SplitWeight(weight.detach()), None if bias is None else bias.detach())

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if _gemm_ok(w.size(1), w.size(0)):
                # dX = dY W on the tcgen05 kernel. The split-fp16 operands need A in the NORMAL fp16 range (activations are O(1),
                # gradients are ~1e-6 and would lose their lo plane): scale by an exact power of two on the device, undo afterwards
                sc = torch.exp2(torch.round(torch.log2(8.0 / gy.abs().amax().clamp_min(1e-30))))
                gx = ops.linear(gy * sc, ops.SplitWeight(w.detach().t().contiguous())) / sc
            else:
                gx = gy @ w
        if ctx.needs_input_grad[1]:
            gw = gy.t() @ x                                                                  # plain library GEMM, [out, in]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb


def linear(x, weight, bias=None):
    return _LinearFn.apply(x, weight, bias)


def transpose_csr(csr):
    """CSR of the reversed edges with the same GCN weights: row j lists the destinations i of the edges j -> i (ascending edge
    position), so that the backward of out = A_hat h is the same fixed-order SpMM kernel applied to d(out)."""
    n = csr.N
    dev = csr.rowptr.device
    e = int(csr.rowptr[n].item())
    counts = (csr.rowptr[1:n + 1] - csr.rowptr[:n]).long()
    dst = torch.repeat_interleave(torch.arange(n, device=dev), counts)
    src = csr.col[:e].long()
    order = torch.sort(src, stable=True).indices
    t_rowptr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    if e:
        t_rowptr[1:] = torch.cumsum(torch.bincount(src, minlength=n), 0).to(torch.int32)
    pad = max(e, 1)
    t_col = torch.zeros(pad, dtype=torch.int32, device=dev)
    t_norm = torch.zeros(pad, dtype=torch.float32, device=dev)
    t_col[:e] = dst[order].to(torch.int32)
    t_norm[:e] = csr.norm[:e][order]
    return ops.CSR(t_rowptr, t_col, None, t_norm, csr.dinv, n, e)


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, addend, bias, csr, csr_t, n_rows):
        ctx.csr, ctx.csr_t, ctx.n_rows = csr, csr_t, n_rows
        ctx.has = (addend is not None, bias is not None)
        ctx.n_add = 0 if addend is None else addend.size(0)
        return ops.spmm_gcn(csr, h.contiguous(), n_rows=n_rows, addend=None if addend is None else addend.contiguous(),
                            bias=None if bias is None else bias.detach())

    @staticmethod
    def backward(ctx, go):
        go = go.contiguous()
        gh = ops.spmm_gcn(ctx.csr_t, go, n_rows=ctx.n_rows) if ctx.needs_input_grad[0] else None
        ga = go[:ctx.n_add] if ctx.has[0] and ctx.needs_input_grad[1] else None
        gb = go.sum(0) if ctx.has[1] and ctx.needs_input_grad[2] else None
        return gh, ga, gb, None, None, None


class _SpatialAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, pos, seg_ptr, seg_of, num_seg, heads, temperature, max_seg):
        qkv = qkv.contiguous()
        out, lse = ops.spatial_attention(qkv, pos, seg_ptr, num_seg, heads, temperature, max_seg=max_seg, want_lse=True)
        ctx.save_for_backward(qkv, out, lse)
        ctx.args = (pos, seg_ptr, seg_of, heads, temperature)
        return out

    @staticmethod
    def backward(ctx, go):
        qkv, out, lse = ctx.saved_tensors
        pos, seg_ptr, seg_of, heads, temperature = ctx.args
        return (ops.spatial_attention_backward(qkv, pos, seg_of, seg_ptr, heads, temperature, out, go, lse=lse),
                None, None, None, None, None, None, None)


# ------------------------------------------------------------------------------------------------ layers
def _norm(m, x, training):
    if isinstance(m, nn.LayerNorm):
        return F.layer_norm(x, (x.size(-1),), m.weight, m.bias, m.eps)
    if isinstance(m, nn.InstanceNorm1d):
        return F.layer_norm(x, (x.size(-1),), None, None, m.eps)
    if isinstance(m, nn.BatchNorm1d):
        return F.batch_norm(x, m.running_mean, m.running_var, m.weight, m.bias, training, m.momentum or 0.0, m.eps)
    return x


class _Level:
    """A graph level for the training path: CSR, its transpose and the (constant) aggregated edge attributes."""

    def __init__(self, csr, agg):
        self.csr, self.csr_t, self.agg = csr, transpose_csr(csr), agg


def _conv(conv, x, lvl, n_rows=None):
    """GraphConvolution (core/graph_layers.py:68-110): A_hat (x W^T) + W_e (sum_e n_e a_e) + b."""
    h = linear(x, conv.node_lin.weight)
    addend = None
    if conv.edge_lin is not None and lvl.agg is not None:
        addend = linear(lvl.agg, conv.edge_lin.weight)
    return _SpmmFn.apply(h, addend, conv.bias, lvl.csr, lvl.csr_t, x.size(0) if n_rows is None else n_rows)


def _dyn_layer(layer, x, lvl, training, n_rows=None):
    """DynamicGraphLayer.forward (core/graph_layers.py:207-247; the dead edge-weight branch D14 omitted)."""
    p = layer.dropout.p
    h1 = F.dropout(F.gelu(_conv(layer.graph_conv1, x, lvl, n_rows)), p, training)
    h2 = F.dropout(F.gelu(_conv(layer.graph_conv2, h1, lvl, n_rows)), p, training)
    out = linear(h2, layer.output_proj.weight, layer.output_proj.bias)
    return _norm(layer.norm1, out + x, training)


def _feature_encoder(fe, x, training):
    act = _ACT[fe.activation_name]
    h = x
    for i in range(fe.num_layers):
        lin, norm, drop = fe.encoder[4 * i], fe.encoder[4 * i + 1], fe.encoder[4 * i + 3]
        h = F.dropout(act(_norm(norm, linear(h, lin.weight, lin.bias), training)), drop.p, training)
    if fe.use_residual:
        h = h + (x if fe.residual_proj is None else linear(x, fe.residual_proj.weight, fe.residual_proj.bias))
    return h


def _graph_encoder(ge, x, lvl, training):
    act = _ACT[ge.activation_name]
    h = x
    for i, layer in enumerate(ge.graph_layers):
        h = _dyn_layer(layer, h, lvl, training)
        tr = ge.transitions[i]
        if not isinstance(tr, nn.Identity):
            h = linear(h, tr.weight, tr.bias)
        h = F.dropout(act(_norm(ge.norm_layers[i], h, training)), ge.dropout.p, training)
    return linear(h, ge.output_proj.weight, ge.output_proj.bias)


def _spatial_attention(sa, x, pos, ctx, training):
    att = sa.attention
    pos = pos.float().contiguous()
    pe = ops.spatial_pe_add(torch.zeros_like(x), pos, ctx.seg_ptr, ctx.num_graphs)        # constant w.r.t. the parameters
    w = torch.cat([att.q_proj.weight, att.k_proj.weight, att.v_proj.weight], 0)
    b = torch.cat([l.bias if l.bias is not None else l.weight.new_zeros(l.weight.size(0)) for l in (att.q_proj, att.k_proj, att.v_proj)])
    qkv = linear(x + pe, w, b)
    seg_of = torch.repeat_interleave(torch.arange(ctx.num_graphs, device=x.device, dtype=torch.int32),
                                     torch.tensor([ctx.seg_host[i + 1] - ctx.seg_host[i] for i in range(ctx.num_graphs)], device=x.device))
    max_seg = max(ctx.seg_host[i + 1] - ctx.seg_host[i] for i in range(ctx.num_graphs))
    if sa.head_dim != 16:
        raise NotImplementedError("the attention backward kernel supports head width 16 (DGDM-Base)")
    o = _SpatialAttnFn.apply(qkv, pos, ctx.seg_ptr, seg_of.contiguous(), ctx.num_graphs, sa.num_heads, sa.temperature, max_seg)
    o = F.dropout(linear(o, att.out_proj.weight, att.out_proj.bias), att.resid_dropout.p, training)
    return _norm(sa.norm, x + o, training)


def _unet(unet, x, ctx, ea, training):
    """GraphUNet.forward (core/graph_layers.py:400-458, P5, D8); the top-k selection itself is not differentiable, the gradient
    flows through x[perm] * score[perm] exactly as in the reference."""
    act = _ACT[unet.act_name]
    csr0 = ctx.csr

    def level(csr):
        return _Level(csr, None if ea is None else ops.edge_attr_aggregate(csr, ea))
    lv = [level(csr0)]
    x = _dyn_layer(unet.down_convs[0], x, lv[0], training)
    xs, perms = [x], []
    for i in range(unet.depth):
        x = _dyn_layer(unet.down_convs[i + 1], act(x), lv[-1], training)
        xs.append(x)
        pool = unet.pools[i]
        l0, l2 = pool.score_net[0], pool.score_net[2]
        s = torch.tanh(linear(F.relu(linear(x, l0.weight, l0.bias)), l2.weight, l2.bias).squeeze(-1))
        k = max(1, int(pool.ratio * x.size(0)))
        perm, node_map = ops.topk_select(s.detach().contiguous(), k)
        perms.append(perm.long())
        lv.append(level(ops.csr_pool(lv[-1].csr, perm, node_map, ea is not None)))
        x = x[perms[-1]] * s[perms[-1]].unsqueeze(-1)
    x = _dyn_layer(unet.bottom_conv, act(x), lv[-1], training)
    for i in range(unet.depth):
        j = unet.depth - 1 - i
        up = torch.zeros(xs[j + 1].size(0), x.size(1), device=x.device, dtype=x.dtype).index_copy(0, perms[j], x)
        x = _dyn_layer(unet.up_convs[i], act(up + xs[j + 1]), lv[j + 1], training, n_rows=xs[j + 1].size(0))      # D8
    return linear(x, unet.final_conv.weight, unet.final_conv.bias)


def _diffusion_loss(model, h, ctx, rng):
    """models/dgdm_model.py:405-445 with P6; D13 (independent noise target) reproduced."""
    dl = model.diffusion_layer
    B, dev = ctx.num_graphs, h.device
    rng = rng or {}
    t = rng.get("timesteps")
    t = torch.randint(0, model.num_diffusion_steps, (B,), device=dev) if t is None else t.to(dev)
    noise = rng.get("noise")
    target = rng.get("target")
    noise = torch.randn_like(h) if noise is None else noise.to(dev)
    target = torch.randn_like(h) if target is None else target.to(dev)
    sizes = torch.tensor([ctx.seg_host[i + 1] - ctx.seg_host[i] for i in range(B)], device=dev)
    t_nodes = torch.repeat_interleave(t, sizes)
    sa, sb = dl._dev_tables(dev)
    xn = sa[t_nodes].unsqueeze(-1) * h + sb[t_nodes].unsqueeze(-1) * noise
    te0, te2 = dl.time_embed[0], dl.time_embed[2]
    emb = dl.get_timestep_embedding(t)                                                     # [B, 128]
    te = linear(F.silu(linear(emb, te0.weight, te0.bias)), te2.weight, te2.bias)           # [B, hidden]
    d0, g1, d4, g5, d8 = (dl.denoise_net[i] for i in (0, 1, 4, 5, 8))
    hh = linear(torch.cat([xn, torch.repeat_interleave(te, sizes, dim=0)], dim=-1), d0.weight, d0.bias)
    hh = F.dropout(F.silu(F.group_norm(hh, g1.num_groups, g1.weight, g1.bias, g1.eps)), dl.denoise_net[3].p, model.training)
    hh = linear(hh, d4.weight, d4.bias)
    hh = F.dropout(F.silu(F.group_norm(hh, g5.num_groups, g5.weight, g5.bias, g5.eps)), dl.denoise_net[7].p, model.training)
    eps = linear(hh, d8.weight, d8.bias)
    losses = [F.mse_loss(eps[ctx.seg_host[i]:ctx.seg_host[i + 1]], target[ctx.seg_host[i]:ctx.seg_host[i + 1]]) for i in range(B)]
    a, b = ctx.seg_host[B - 1], ctx.seg_host[B]
    return {"diffusion_loss": torch.stack(losses).mean(), "noisy_embeddings": xn[a:b].unsqueeze(0)}


def _global_pool(model, h, ctx):
    gp = model.global_pool
    if not hasattr(gp, "global_token"):
        return gp(h, None, ctx=ctx)
    att = gp.attention
    tok = gp.global_token.reshape(1, -1)
    q = linear(tok, att.q_proj.weight, att.q_proj.bias)                                    # [1, D]
    k = linear(h, att.k_proj.weight, att.k_proj.bias)
    v = linear(h, att.v_proj.weight, att.v_proj.bias)
    H, hd = att.num_heads, att.head_dim
    outs = []
    for g in range(ctx.num_graphs):                                                        # one query per graph: row-wise torch ops
        a, b = ctx.seg_host[g], ctx.seg_host[g + 1]
        if b == a:
            outs.append(h.new_zeros(1, H * hd))
            continue
        s = (k[a:b].view(-1, H, hd) * q.view(1, H, hd)).sum(-1) / hd ** 0.5                # [n, H]
        w = torch.softmax(s, dim=0)
        outs.append((w.unsqueeze(-1) * v[a:b].view(-1, H, hd)).sum(0).reshape(1, -1))
    return linear(torch.cat(outs, 0), att.out_proj.weight, att.out_proj.bias)


def train_forward(model, data, mode="pretrain", return_embeddings=False, rng=None):
    """Differentiable DGDMModel.forward (single GPU, node-sorted batches). Returns the same output dict; losses carry a graph."""
    from .core import _edge_attr_or_none
    x = data.x.contiguous()
    ctx = get_context(data)
    pos = getattr(data, "pos", None)
    training = model.training
    ea = _edge_attr_or_none(getattr(data, "edge_attr", None), 32)
    if getattr(data, "batch", None) is not None and not ctx.sorted:
        raise NotImplementedError("the training path needs a batch vector sorted by graph id")
    lvl0 = _Level(ctx.csr, None if ea is None else ops.edge_attr_aggregate(ctx.csr, ea))
    h = _feature_encoder(model.feature_encoder, x, training)
    h = _graph_encoder(model.graph_encoder, h, lvl0, training)
    if model.spatial_attention is not None and pos is not None:
        h = _spatial_attention(model.spatial_attention, h, pos, ctx, training)
    if model.hierarchical_processor is not None:
        h = _unet(model.hierarchical_processor, h, ctx, ea, training)
    out = {}
    if mode == "pretrain":
        out.update(_diffusion_loss(model, h, ctx, rng))
    g = _global_pool(model, h, ctx)
    if model.classification_head is not None and mode in ("inference", "finetune"):
        out["classification_logits"] = model.classification_head(g)
        out["classification_probs"] = F.softmax(out["classification_logits"], dim=-1)
    if model.regression_head is not None and mode in ("inference", "finetune"):
        out["regression_outputs"] = model.regression_head(g)
    out["graph_embedding"] = g
    if return_embeddings:
        out["node_embeddings"] = h
    return out


def allreduce_gradients(model, group=None, bucket_bytes=64 << 20):
    """Slide-batch data parallelism (config 5): bucketed NCCL all-reduce (mean) of the gradients after backward."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return
    params = [p for p in model.parameters() if p.grad is not None]
    bucket, size = [], 0

    def flush():
        if not bucket:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        dist.all_reduce(flat, group=group)
        flat /= world
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    for p in params:
        bucket.append(p)
        size += p.grad.numel() * 4
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()
