"""TEST INFRASTRUCTURE — CPU restatement (plain torch fp32, functional, state_dict in / tensors out) of the
reference's DGDM forward hot path with the patch set P1-P7 of SURVEY.md §8c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import it; the product
package never does.

Pinning: the reference holds no golden vectors for this path, so this restatement is pinned
against the reference itself, run in the build container through oracle/ref_loader.py:
tests/test_oracle_vs_reference.py (live, skipped where /root/reference is absent) and the committed
fixtures under tests/golden/ produced by oracle/make_golden.py.

All file:line citations are into the reference tree (dgdm_histopath/...).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- helpers
def _lin(sd, p, x, bias=True):
    b = sd.get(p + ".bias") if bias else None
    return F.linear(x, sd[p + ".weight"], b)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.size(-1),), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _norm(sd, p, x, kind="layer"):
    """encoders.py:93-102 / :217-226 `_get_norm_layer` in eval mode: LayerNorm | BatchNorm1d (running statistics) | InstanceNorm1d
    (a 2-D [N, C] input is an unbatched (channels = N, length = C) signal: every row is normalised over its C entries, no affine)."""
    if kind == "layer":
        return _ln(sd, p, x)
    if kind == "batch":
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)
    if kind == "instance":
        return F.layer_norm(x, (x.size(-1),), None, None, 1e-5)
    return x


def _act(name):
    return {"gelu": F.gelu, "relu": F.relu, "elu": F.elu}[name]


def pad_edge_attr(edge_attr, num_edges, width, like):
    """P5/P7 + encoders.py:258-261: missing edge_attr is zeros(E, width); narrower attrs are zero-padded."""
    if edge_attr is None:
        return like.new_zeros(num_edges, width)
    if edge_attr.size(1) < width:
        return torch.cat([edge_attr, edge_attr.new_zeros(edge_attr.size(0), width - edge_attr.size(1))], 1)
    return edge_attr


# ----------------------------------------------------------------------------- message passing
def gcn_norm(edge_index, num_nodes, dtype=torch.float32):
    """core/graph_layers.py:76-84 with PyG add_self_loops/degree semantics (loops appended after the edges)."""
    dev = edge_index.device          # device / dtype follow the inputs: the tests also run this restatement in float64 on the GPU
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=dev)
    ei = torch.cat([edge_index, torch.stack([loops, loops])], dim=1)
    row, col = ei[0], ei[1]
    deg = torch.zeros(num_nodes, dtype=dtype, device=dev).index_add_(0, col, torch.ones(col.numel(), dtype=dtype, device=dev))
    dinv = deg.pow(-0.5)
    dinv[dinv == float("inf")] = 0
    return ei, dinv[row] * dinv[col]


def graph_convolution(sd, p, x, edge_index, edge_attr=None):
    """GraphConvolution.forward/message, core/graph_layers.py:68-110 (+P2: zero attr rows for the loops)."""
    n = x.size(0)
    ei, norm = gcn_norm(edge_index, n, x.dtype)
    h = F.linear(x, sd[p + ".node_lin.weight"])
    msg = h.index_select(0, ei[0])
    if edge_attr is not None and (p + ".edge_lin.weight") in sd:
        ea = torch.cat([edge_attr, edge_attr.new_zeros(n, edge_attr.size(1))], 0)
        msg = msg + F.linear(ea, sd[p + ".edge_lin.weight"])
    msg = norm.view(-1, 1) * msg
    out = torch.zeros(n, h.size(1), dtype=h.dtype, device=h.device).index_add_(0, ei[1], msg)
    if (p + ".bias") in sd:
        out = out + sd[p + ".bias"]
    return out


def segment_softmax(src, index, num_nodes):
    """torch_geometric.utils.softmax semantics: exp(src - segmax) / (segsum + 1e-16)."""
    idx = index.view(-1, 1).expand_as(src)
    smax = torch.full((num_nodes, src.size(1)), float("-inf")).scatter_reduce(0, idx, src, "amax", include_self=True)
    e = (src - smax.gather(0, idx)).exp()
    ssum = torch.zeros(num_nodes, src.size(1)).scatter_add_(0, idx, e)
    return e / (ssum.gather(0, idx) + 1e-16)


def compute_dynamic_edges(sd, p, x, edge_index, edge_attr, num_heads):
    """DynamicGraphLayer.compute_dynamic_edges, core/graph_layers.py:160-205."""
    row, col = edge_index[0], edge_index[1]
    qkv = _lin(sd, p + ".node_to_qkv", x)
    hid = qkv.size(1) // 3
    hd = hid // num_heads
    q = qkv[:, :hid].reshape(-1, num_heads, hd)
    k = qkv[:, hid:2 * hid].reshape(-1, num_heads, hd)
    ek = _lin(sd, p + ".edge_to_key", edge_attr).view(-1, num_heads, hd)
    scores = (q[row] * (k[col] + ek)).sum(-1) / math.sqrt(hd)
    return segment_softmax(scores, col, x.size(0))


def dynamic_graph_layer(sd, p, x, edge_index, edge_attr):
    """DynamicGraphLayer.forward, core/graph_layers.py:207-247 (eval: dropout = identity; D14 dead branch omitted)."""
    h1 = F.gelu(graph_convolution(sd, p + ".graph_conv1", x, edge_index, edge_attr))
    h2 = F.gelu(graph_convolution(sd, p + ".graph_conv2", h1, edge_index, edge_attr))
    return _ln(sd, p + ".norm1", _lin(sd, p + ".output_proj", h2) + x)


def adaptive_graph_pooling(sd, p, x, edge_index, edge_attr=None, ratio=0.5):
    """AdaptiveGraphPooling.forward, core/graph_layers.py:285-329 (tanh scores, batch ignored — D16)."""
    s = _lin(sd, p + ".score_net.2", F.relu(_lin(sd, p + ".score_net.0", x))).squeeze(-1)
    s = torch.tanh(s)
    k = max(1, int(ratio * x.size(0)))
    _, top = torch.topk(s, k, sorted=False)
    mask = torch.zeros_like(s, dtype=torch.bool)
    mask[top] = True
    perm = mask.nonzero(as_tuple=False).squeeze(-1)
    px = x[perm] * s[perm].unsqueeze(-1)
    node_map = torch.full((x.size(0),), -1, dtype=torch.long, device=x.device)
    node_map[perm] = torch.arange(perm.size(0), device=x.device)
    emask = (node_map[edge_index[0]] >= 0) & (node_map[edge_index[1]] >= 0)
    pei = node_map[edge_index[:, emask]]
    pea = edge_attr[emask] if edge_attr is not None else None
    return px, pei, pea, perm, s


def graph_unet(sd, p, x, edge_index, edge_attr=None, depth=3, act="relu", return_perms=False, scores_out=None):
    """GraphUNet.forward, core/graph_layers.py:400-458 with P5 (edge_dim 32, up_convs on hidden); D8 kept.
    scores_out (list): receives the tanh pooling scores of every level (for the top-k flip report)."""
    a = _act(act)
    # None stays None below: edge_lin has no bias, so zeros(E,32) (P5) contribute exactly nothing
    edge_attr = pad_edge_attr(edge_attr, edge_index.size(1), 32, x) if edge_attr is not None else None
    xs, eis, eas, perms = [], [edge_index], [edge_attr], []
    x = dynamic_graph_layer(sd, p + ".down_convs.0", x, edge_index, edge_attr)
    xs.append(x)
    for i in range(depth):
        x = dynamic_graph_layer(sd, f"{p}.down_convs.{i + 1}", a(x), eis[-1], eas[-1])
        xs.append(x)
        x, ei, ea, perm, sc = adaptive_graph_pooling(sd, f"{p}.pools.{i}", x, eis[-1], eas[-1])
        eis.append(ei); eas.append(ea); perms.append(perm)
        if scores_out is not None:
            scores_out.append(sc)
    x = dynamic_graph_layer(sd, p + ".bottom_conv", a(x), eis[-1], eas[-1])
    for i in range(depth):
        j = depth - 1 - i
        up = torch.zeros(xs[j + 1].size(0), x.size(1), dtype=x.dtype, device=x.device)
        up[perms[j]] = x
        x = a(up + xs[j + 1])
        x = dynamic_graph_layer(sd, f"{p}.up_convs.{i}", x, eis[j + 1], eas[j + 1])   # D8: pooled edges, level-j nodes
    x = _lin(sd, p + ".final_conv", x)
    return (x, perms) if return_perms else x


# ----------------------------------------------------------------------------- encoders
def feature_encoder(sd, p, x, act="gelu", norm="layer"):
    """FeatureEncoder.forward, models/encoders.py:104-124 (2 layers, eval)."""
    a = _act(act)
    h = a(_norm(sd, p + ".encoder.1", _lin(sd, p + ".encoder.0", x), norm))
    h = a(_norm(sd, p + ".encoder.5", _lin(sd, p + ".encoder.4", h), norm))
    if (p + ".residual_proj.weight") in sd:
        return h + _lin(sd, p + ".residual_proj", x)
    return h + x


def graph_encoder(sd, p, x, edge_index, edge_attr, num_layers, act="gelu", return_layers=False, norm="layer"):
    """GraphEncoder.forward, models/encoders.py:228-280 with P1/P3."""
    a = _act(act)
    ea = pad_edge_attr(edge_attr, edge_index.size(1), 32, x)
    h, outs = x, []
    for i in range(num_layers):
        h = dynamic_graph_layer(sd, f"{p}.graph_layers.{i}", h, edge_index, ea)
        if f"{p}.transitions.{i}.weight" in sd:
            h = _lin(sd, f"{p}.transitions.{i}", h)
        h = a(_norm(sd, f"{p}.norm_layers.{i}", h, norm))
        outs.append(h)
    out = _lin(sd, p + ".output_proj", h)
    return (out, outs) if return_layers else out


# ----------------------------------------------------------------------------- attention
def multi_head_attention(sd, p, query, key=None, value=None, num_heads=8, attn_mask=None, need_weights=True,
                         key_padding_mask=None, average_attn_weights=True):
    """MultiHeadAttention.forward, core/attention.py:73-181 (batch_first, eval). Inputs [B, L, D]. Masks follow the reference
    literally: bool attn_mask -> masked_fill(-inf) (:140), float -> `scores += mask` with torch broadcasting (:142),
    key_padding_mask [B, Lk] -> -inf (:145-150); weights head-averaged (:170) or [B*H, Lq, Lk] (:172)."""
    key = query if key is None else key
    value = key if value is None else value
    b, lq, d = query.shape
    hd = d // num_heads
    q = _lin(sd, p + ".q_proj", query).view(b, lq, num_heads, hd).transpose(1, 2)
    k = _lin(sd, p + ".k_proj", key).view(b, -1, num_heads, hd).transpose(1, 2)
    v = _lin(sd, p + ".v_proj", value).view(b, -1, num_heads, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)
    if attn_mask is not None:
        s = s.masked_fill(attn_mask, float("-inf")) if attn_mask.dtype == torch.bool else s + attn_mask
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(2), float("-inf"))
    w = F.softmax(s, dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).contiguous().view(b, lq, d)
    o = _lin(sd, p + ".out_proj", o)
    if not need_weights:
        return o, None
    return o, (w.mean(dim=1) if average_attn_weights else w.reshape(b * num_heads, lq, -1))


def spatial_positional_encoding(pos, embed_dim):
    """SpatialAttention.get_positional_encoding, core/attention.py:225-259. pos [..., 2] -> [..., embed_dim]; the min / max
    are taken over the WHOLE tensor (:233-236)."""
    pn = pos.float()
    if pn.numel() > 0:
        pn = (pn - pn.min()) / (pn.max() - pn.min() + 1e-8)
    pe = torch.zeros(tuple(pos.shape[:-1]) + (embed_dim,))
    div = torch.exp(torch.arange(0, embed_dim // 2, 2) * -(math.log(10000.0) / (embed_dim // 2)))
    pe[..., 0::4] = torch.sin(pn[..., 0:1] * div)
    pe[..., 1::4] = torch.cos(pn[..., 0:1] * div)
    pe[..., 2::4] = torch.sin(pn[..., 1:2] * div)
    pe[..., 3::4] = torch.cos(pn[..., 1:2] * div)
    return pe


def spatial_attention_batched(sd, p, x, pos, num_heads=8, temperature=1.0, mask=None):
    """SpatialAttention.forward with its public signature, core/attention.py:285-327: x [B,n,D], pos [B,n,2]. The bias
    [B,n,n] is added to scores [B,H,n,n] by plain torch broadcasting — defined only for B == 1 or (accidentally) B == H."""
    pe = spatial_positional_encoding(pos, x.size(-1))
    bias = -torch.norm(pos.unsqueeze(2) - pos.unsqueeze(1), dim=-1) / temperature
    if mask is not None:
        bias = mask + bias
    o, w = multi_head_attention(sd, p + ".attention", x + pe, num_heads=num_heads, attn_mask=bias, need_weights=True)
    return _ln(sd, p + ".norm", x + o), w


def spatial_attention(sd, p, x, pos, num_heads=8, temperature=1.0, need_weights=False):
    """SpatialAttention.forward for ONE graph, core/attention.py:285-327. x [N,D], pos [N,2]."""
    pe = spatial_positional_encoding(pos, x.size(1))
    bias = -torch.norm(pos.unsqueeze(1) - pos.unsqueeze(0), dim=-1) / temperature
    o, w = multi_head_attention(sd, p + ".attention", (x + pe).unsqueeze(0), num_heads=num_heads,
                                attn_mask=bias.unsqueeze(0), need_weights=need_weights)
    out = _ln(sd, p + ".norm", x + o.squeeze(0))
    return out, (w.squeeze(0) if w is not None else None)


def global_attention_pool(sd, p, x, batch, num_heads=8):
    """GlobalAttentionPool.forward, models/dgdm_model.py:588-615."""
    tok = sd[p + ".global_token"]
    if batch is None:
        o, _ = multi_head_attention(sd, p + ".attention", tok, x.unsqueeze(0), x.unsqueeze(0), num_heads, need_weights=False)
        return o.squeeze(1)
    nb = int(batch.max()) + 1
    out = torch.zeros(nb, x.size(1))
    for i in range(nb):
        m = batch == i
        if m.any():
            xi = x[m].unsqueeze(0)
            o, _ = multi_head_attention(sd, p + ".attention", tok, xi, xi, num_heads, need_weights=False)
            out[i] = o.squeeze(1)
    return out


# ----------------------------------------------------------------------------- diffusion
def diffusion_schedule(num_timesteps, schedule="cosine", beta_start=1e-4, beta_end=0.02):
    """DiffusionScheduler.__init__, core/diffusion.py:16-61. Returns dict of [T] fp32 tables."""
    if schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, num_timesteps)
    elif schedule == "cosine":
        x = torch.linspace(0, num_timesteps, num_timesteps + 1)
        ac = torch.cos(((x / num_timesteps) + 0.008) / 1.008 * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    elif schedule == "sigmoid":
        betas = torch.sigmoid(torch.linspace(-6, 6, num_timesteps)) * (beta_end - beta_start) + beta_start
    else:
        raise ValueError(f"Unknown schedule: {schedule}")
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
    return dict(betas=betas, alphas=alphas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev,
                posterior_variance=betas * (1.0 - ac_prev) / (1.0 - ac))


def timestep_embedding(t, dim=128):
    """DiffusionLayer.get_timestep_embedding, core/diffusion.py:112-121."""
    half = dim // 2
    f = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    e = t.float()[:, None] * f[None, :]
    return torch.cat([torch.sin(e), torch.cos(e)], dim=1)


def diffusion_add_noise(sched, x0, noise, t):
    """DiffusionLayer.add_noise, core/diffusion.py:123-145. x0 [N,D], t [N] long."""
    a = torch.sqrt(sched["alphas_cumprod"][t]).unsqueeze(-1)
    b = torch.sqrt(1.0 - sched["alphas_cumprod"][t]).unsqueeze(-1)
    return a * x0 + b * noise


def diffusion_predict_noise(sd, p, x_noisy, t, condition=None):
    """DiffusionLayer.predict_noise, core/diffusion.py:147-172, 2-D input (P6), eval (dropout off); condition: t_emb + condition_net(c)."""
    te = _lin(sd, p + ".time_embed.2", F.silu(_lin(sd, p + ".time_embed.0", timestep_embedding(t))))
    if condition is not None and (p + ".condition_net.weight") in sd:
        te = te + _lin(sd, p + ".condition_net", condition)
    h = _lin(sd, p + ".denoise_net.0", torch.cat([x_noisy, te], dim=-1))
    h = F.silu(F.group_norm(h, 8, sd[p + ".denoise_net.1.weight"], sd[p + ".denoise_net.1.bias"], 1e-5))
    h = _lin(sd, p + ".denoise_net.4", h)
    h = F.silu(F.group_norm(h, 8, sd[p + ".denoise_net.5.weight"], sd[p + ".denoise_net.5.bias"], 1e-5))
    return _lin(sd, p + ".denoise_net.8", h)


def diffusion_forward(sd, p, sched, x0, t, noise):
    """DiffusionLayer.forward, core/diffusion.py:174-212 with injected noise. t [N] long."""
    xn = diffusion_add_noise(sched, x0, noise, t)
    return xn, diffusion_predict_noise(sd, p, xn, t)


def diffusion_sample(sd, p, sched, init, step_noise, num_inference_steps=50):
    """DiffusionLayer.sample, core/diffusion.py:214-275 with injected initial sample and per-step noise."""
    T = sched["betas"].numel()
    x = init
    ts = torch.linspace(T - 1, 0, num_inference_steps, dtype=torch.long)
    for i, t in enumerate(ts):
        tb = torch.full((x.size(0),), int(t), dtype=torch.long)
        eps = diffusion_predict_noise(sd, p, x, tb)
        ac = sched["alphas_cumprod"][t]
        x0 = (x - torch.sqrt(1 - ac) * eps) / torch.sqrt(ac)
        if i < len(ts) - 1:
            x = torch.sqrt(sched["alphas"][t]) * x0 + torch.sqrt(sched["posterior_variance"][t]) * step_noise[i]
        else:
            x = x0
    return x


# ----------------------------------------------------------------------------- heads
def classification_head(sd, p, x, act="gelu"):
    """ClassificationHead.forward, models/decoders.py:89-99 — Linear, BatchNorm1d(eval), act, Linear."""
    a = _act(act)
    h = _lin(sd, p + ".classifier.0", x)
    if (p + ".classifier.1.running_mean") in sd:
        h = F.batch_norm(h, sd[p + ".classifier.1.running_mean"], sd[p + ".classifier.1.running_var"],
                         sd[p + ".classifier.1.weight"], sd[p + ".classifier.1.bias"], False, 0.0, 1e-5)
        return _lin(sd, p + ".classifier.4", a(h))
    return _lin(sd, p + ".classifier.3", a(h))


# ----------------------------------------------------------------------------- model
def apply_entity_masking(x, mask_indices, mask_token):
    """DGDMModel._apply_entity_masking, models/dgdm_model.py:482-506 with the draws injected (randperm(n)[:k], randn(F))."""
    out = x.clone()
    if mask_indices is not None and mask_indices.numel() > 0:
        node_mask = torch.zeros(x.size(0), dtype=torch.bool)
        node_mask[mask_indices] = True
        out[node_mask] = mask_token
    return out


def pretrain_step(sd, cfg, x, edge_index, edge_attr=None, batch=None, pos=None, rng=None):
    """DGDMModel.pretrain_step, models/dgdm_model.py:447-480: mask, forward(mode='pretrain'); the reconstruction branch is
    dead (D12) so total_pretrain_loss = diffusion_loss."""
    xm = apply_entity_masking(x, rng.get("mask_indices"), rng.get("mask_token"))
    out = dgdm_forward(sd, cfg, xm, edge_index, edge_attr, batch, pos, mode="pretrain", rng=rng)
    out["total_pretrain_loss"] = out["diffusion_loss"]
    return out


def dgdm_forward(sd, cfg, x, edge_index, edge_attr=None, batch=None, pos=None, mode="inference",
                 rng=None, return_stages=False):
    """DGDMModel.forward + _forward_continue, models/dgdm_model.py:271-403 (eval; P1-P7).

    cfg: dict(hidden_dims, attention_heads, graph_layers, activation, use_spatial_attention, use_hierarchical,
              num_diffusion_steps, diffusion_schedule).  rng (pretrain): dict(timesteps [B] long,
    noise: list of [N_i, D], target: list of [N_i, D]) — injected, see SURVEY.md §8c "RNG".
    """
    heads = cfg.get("attention_heads", 8)
    act = cfg.get("activation", "gelu")
    st = {}
    norm = cfg.get("normalization", "layer")
    h = feature_encoder(sd, "feature_encoder", x, act, norm)
    st["feature_encoder"] = h
    h = graph_encoder(sd, "graph_encoder", h, edge_index, edge_attr, cfg.get("graph_layers", 4), act, norm=norm)
    st["graph_encoder"] = h
    if cfg.get("use_spatial_attention", True) and pos is not None:
        if batch is not None:
            parts = []
            for i in range(int(batch.max()) + 1):
                m = batch == i
                o, _ = spatial_attention(sd, "spatial_attention", h[m], pos[m], heads)
                parts.append(o)
            h = torch.cat(parts, 0)
        else:
            h, _ = spatial_attention(sd, "spatial_attention", h, pos, heads)
    st["spatial_attention"] = h
    if cfg.get("use_hierarchical", True):
        h, perms = graph_unet(sd, "hierarchical_processor", h, edge_index, edge_attr, return_perms=True)
        st["perms"] = perms
    out = {}
    if mode == "pretrain":
        sched = diffusion_schedule(cfg.get("num_diffusion_steps", 10), cfg.get("diffusion_schedule", "cosine"))
        losses, xn = [], None
        groups = [torch.ones(h.size(0), dtype=torch.bool)] if batch is None else \
            [batch == i for i in range(int(batch.max()) + 1)]
        for i, m in enumerate(groups):
            hi = h[m]
            t = rng["timesteps"][i:i + 1].repeat(hi.size(0))
            xn, eps = diffusion_forward(sd, "diffusion_layer", sched, hi, t, rng["noise"][i])
            losses.append(F.mse_loss(eps, rng["target"][i]))                    # D13: independent target noise
        out["diffusion_loss"] = torch.stack(losses).mean()
        out["noisy_embeddings"] = xn.unsqueeze(0)                               # last graph only (dgdm_model.py:442-445)
    pooling = cfg.get("pooling", "attention")
    if pooling == "attention":
        g = global_attention_pool(sd, "global_pool", h, batch, heads)
    else:
        # GlobalMeanPool / GlobalMaxPool / GlobalSet2SetPool (a mean in disguise), models/dgdm_model.py:552-585,618-642
        groups = [torch.ones(h.size(0), dtype=torch.bool)] if batch is None else [batch == i for i in range(int(batch.max()) + 1)]
        g = torch.stack([h[m].max(dim=0)[0] if pooling == "max" else h[m].mean(dim=0) for m in groups])
    out["graph_embedding"] = g
    out["node_embeddings"] = h
    if mode in ("inference", "finetune") and "classification_head.classifier.0.weight" in sd:
        logits = classification_head(sd, "classification_head", g, act)
        out["classification_logits"] = logits
        out["classification_probs"] = F.softmax(logits, dim=-1)
    if return_stages:
        out["stages"] = st
    return out
