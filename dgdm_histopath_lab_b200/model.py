"""DGDMModel — drop-in for the reference's models/dgdm_model.py:37-715 forward path on B200.

Same constructor signature, attributes, state_dict key layout (plus `graph_encoder.transitions.*` of patch P3),
output dict keys and error types; every node/edge-scale computation runs in the sm_100a kernels of csrc/.
The reference forward does not execute as shipped; the behaviour reproduced here is the
reference + patch set P1-P7 that oracle/dgdm_oracle.py restates and tests/ pin.
"""
import logging
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .core import (DiffusionLayer, FeatureEncoder, GraphEncoder, GraphUNet, MultiHeadAttention, SpatialAttention,
                   split_weight, _require_cuda)
from .errors import (ModelConfigurationError, ModelInferenceError, ValidationError, validate_boolean, validate_enum,
                     validate_integer, validate_numeric)
from .graph import GRAPH_TYPES, Data, get_context

_ACTS = {"relu": nn.ReLU, "gelu": nn.GELU, "elu": nn.ELU}


class ClassificationHead(nn.Module):
    """models/decoders.py:14-99 — graph-level MLP on [B, D]; tiny, stays in torch."""

    def __init__(self, input_dim, num_classes, hidden_dims=None, dropout=0.1, activation="gelu", use_batch_norm=True,
                 class_weights=None, label_smoothing=0.0):
        super().__init__()
        self.num_classes, self.label_smoothing = num_classes, label_smoothing
        hidden_dims = hidden_dims if hidden_dims is not None else [input_dim // 2]
        act = _ACTS.get(activation, nn.ReLU)()
        dims, layers = [input_dim] + list(hidden_dims), []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if use_batch_norm:
                layers.append(nn.BatchNorm1d(dims[i + 1]))
            layers += [act, nn.Dropout(dropout)]
        layers.append(nn.Linear(dims[-1], num_classes))
        self.classifier = nn.Sequential(*layers)
        if class_weights is not None:
            self.register_buffer("class_weights", class_weights)
        else:
            self.class_weights = None

    def forward(self, x):
        return self.classifier(x)


class RegressionHead(nn.Module):
    """models/decoders.py:150-261."""

    def __init__(self, input_dim, num_targets, hidden_dims=None, dropout=0.1, activation="gelu", use_batch_norm=True,
                 output_activation=None, predict_uncertainty=False):
        super().__init__()
        self.num_targets, self.predict_uncertainty = num_targets, predict_uncertainty
        hidden_dims = hidden_dims if hidden_dims is not None else [input_dim // 2]
        act = _ACTS.get(activation, nn.ReLU)()
        self.output_activation = {"sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "softplus": nn.Softplus}.get(output_activation, nn.Identity)()
        dims, layers = [input_dim] + list(hidden_dims), []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if use_batch_norm:
                layers.append(nn.BatchNorm1d(dims[i + 1]))
            layers += [act, nn.Dropout(dropout)]
        self.feature_layers = nn.Sequential(*layers)
        self.mean_head = nn.Linear(dims[-1], num_targets)
        self.var_head = nn.Linear(dims[-1], num_targets) if predict_uncertainty else None

    def forward(self, x):
        f = self.feature_layers(x)
        mean = self.output_activation(self.mean_head(f))
        if self.predict_uncertainty:
            lv = self.var_head(f)
            return {"mean": mean, "var": torch.exp(lv), "log_var": lv}
        return mean


def _segment_reduce(x, ctx, op):
    out = torch.zeros(ctx.num_graphs, x.size(1), device=x.device, dtype=x.dtype)
    for g in range(ctx.num_graphs):
        a, b = ctx.seg_host[g], ctx.seg_host[g + 1]
        if b > a:
            out[g] = x[a:b].mean(dim=0) if op == "mean" else x[a:b].max(dim=0)[0]
    return out


class GlobalMeanPool(nn.Module):
    """models/dgdm_model.py:552-567 (non-benchmarked option; torch reduction over node-sorted segments)."""

    def forward(self, x, batch=None, *, ctx=None):
        if batch is None and ctx is None:
            return x.mean(dim=0, keepdim=True)
        return _segment_reduce(x, ctx, "mean")


class GlobalMaxPool(nn.Module):
    """models/dgdm_model.py:570-585."""

    def forward(self, x, batch=None, *, ctx=None):
        if batch is None and ctx is None:
            return x.max(dim=0, keepdim=True)[0]
        return _segment_reduce(x, ctx, "max")


class GlobalSet2SetPool(nn.Module):
    """models/dgdm_model.py:618-642 — a mean in disguise; the LSTM parameters exist but are never used."""

    def __init__(self, hidden_dim, num_layers=2):
        super().__init__()
        self.hidden_dim, self.num_layers = hidden_dim, num_layers
        self.lstm = nn.LSTM(hidden_dim, hidden_dim, num_layers, batch_first=True)

    def forward(self, x, batch=None, *, ctx=None):
        if batch is None and ctx is None:
            return x.mean(dim=0, keepdim=True)
        return _segment_reduce(x, ctx, "mean")


class GlobalAttentionPool(nn.Module):
    """models/dgdm_model.py:588-615: per graph MHA(global_token, X_g, X_g) — one query, so a segmented softmax-reduction
    kernel rather than a GEMM."""

    def __init__(self, hidden_dim, num_heads=8):
        super().__init__()
        self.attention = MultiHeadAttention(hidden_dim, num_heads)
        self.global_token = nn.Parameter(torch.randn(1, 1, hidden_dim))

    @torch.no_grad()
    def forward(self, x, batch=None, *, ctx=None, seg_ptr=None, num_seg=None, part=None):
        """part (parallel.Part): x holds the owned rows; the k|v projections are computed on them and all-gathered (N x 2D
        fp32), then every rank reduces the whole batch with the same kernel and order as a single GPU."""
        _require_cuda(x)
        att = self.attention
        if seg_ptr is None:
            if ctx is not None:
                seg_ptr, num_seg = ctx.seg_ptr, ctx.num_graphs
            elif batch is None:
                seg_ptr, num_seg = torch.tensor([0, x.size(0)], dtype=torch.int32, device=x.device), 1
            else:
                counts = torch.bincount(batch)
                seg_ptr = F.pad(torch.cumsum(counts, 0), (1, 0)).to(torch.int32)
                num_seg = int(counts.numel())
        tok = self.global_token.detach().reshape(1, -1).contiguous()
        q = ops.linear(tok, split_weight(att.q_proj.weight), att.q_proj.bias.detach() if att.q_proj.bias is not None else None)
        w, b = att.kv_weight()
        kv = ops.linear(x.contiguous(), w, b)
        if part is not None:
            kv = part.gather(kv)
        pooled = ops.attention_pool(q.reshape(-1), kv, seg_ptr, num_seg, att.num_heads)
        return ops.linear(pooled, split_weight(att.out_proj.weight), att.out_proj.bias.detach() if att.out_proj.bias is not None else None)


class DGDMModel(nn.Module):
    """Dynamic Graph Diffusion Model — same public API as the reference class (models/dgdm_model.py:45-61,271-277,447,527)."""
    _warned_no_grad = False

    def __init__(self, node_features=768, hidden_dims=[512, 256, 128], num_diffusion_steps=10, attention_heads=8, dropout=0.1,
                 graph_layers=4, use_spatial_attention=True, use_hierarchical=True, diffusion_schedule="cosine",
                 activation="gelu", normalization="layer", pooling="attention", num_classes=None, regression_targets=0):
        super().__init__()
        self.logger = logging.getLogger(__name__)
        try:
            self._validate_configuration(node_features, hidden_dims, num_diffusion_steps, attention_heads, dropout, graph_layers,
                                         diffusion_schedule, activation, normalization, pooling, num_classes, regression_targets)
        except Exception as e:
            raise ModelConfigurationError(f"Invalid model configuration: {e}")
        self.node_features, self.hidden_dims = node_features, hidden_dims
        self.num_diffusion_steps, self.attention_heads, self.dropout = num_diffusion_steps, attention_heads, dropout
        self.use_spatial_attention, self.use_hierarchical, self.pooling = use_spatial_attention, use_hierarchical, pooling
        self.num_classes, self.regression_targets = num_classes, regression_targets
        try:
            self.feature_encoder = FeatureEncoder(node_features, hidden_dims[0], dropout=dropout, activation=activation,
                                                  normalization=normalization)
            self.graph_encoder = GraphEncoder(hidden_dims[0], hidden_dims, num_layers=graph_layers, attention_heads=attention_heads,
                                              dropout=dropout, activation=activation, normalization=normalization)
        except (NotImplementedError, ValueError) as e:
            raise ModelConfigurationError(f"Invalid model configuration: {e}")
        self.diffusion_layer = DiffusionLayer(hidden_dims[-1], hidden_dims[-1] * 2, num_timesteps=num_diffusion_steps,
                                              schedule=diffusion_schedule)
        self.spatial_attention = SpatialAttention(hidden_dims[-1], attention_heads, dropout=dropout) if use_spatial_attention else None
        self.hierarchical_processor = GraphUNet(hidden_dims[-1], hidden_dims[-1], hidden_dims[-1], depth=3) if use_hierarchical else None
        self.global_pool = self._create_pooling_layer(pooling, hidden_dims[-1], attention_heads)
        self.classification_head = None
        self.regression_head = None
        if num_classes is not None:
            self.classification_head = ClassificationHead(hidden_dims[-1], num_classes, [hidden_dims[-1] // 2], dropout, activation)
        if regression_targets > 0:
            self.regression_head = RegressionHead(hidden_dims[-1], regression_targets, [hidden_dims[-1] // 2], dropout, activation)
        self.apply(self._init_weights)
        self._comm = None

    # -------------------------------------------------------------------------------------- multi-GPU
    def set_partition(self, comm):
        """Node-partitioned execution over the ranks of `comm` (parallel.TorchDistComm for one process per GPU over NCCL;
        parallel.ThreadComm for in-process test ranks); None or a 1-rank comm restores single-GPU execution.
        In this mode `data.edge_index`, `data.pos`, `data.batch` describe the WHOLE graph batch on every rank and `data.x`
        holds either all rows or only the rows this rank owns (then `data.num_nodes` gives N); `node_embeddings` in the
        output are the owned rows, `node_range` their (start, stop). See parallel.py for the design."""
        self._comm = comm if comm is not None and comm.world > 1 else None

    def set_distributed(self, rank, world, group=None, max_nodes=None, fused=True):
        """Convenience for torchrun launches: partition over the default (or given) torch.distributed group. With
        max_nodes (the largest N that will be seen) and CUDA peer access the halo exchange is fused into the GEMM epilogue
        (parallel.SymmComm); otherwise it is an NCCL all-gather per propagate."""
        if not world or world <= 1:
            return self.set_partition(None)
        from .parallel import SymmComm, TorchDistComm
        comm = None
        if fused and max_nodes and next(self.parameters()).is_cuda:
            width = max([self.hidden_dims[0]] + list(self.hidden_dims))
            try:
                comm = SymmComm(int(max_nodes) * min(width, 512), next(self.parameters()).device, group)
            except Exception as e:      # no P2P / symmetric memory on this box: NCCL path (still correct, just slower)
                warnings.warn(f"symmetric-memory halo exchange unavailable ({e}); using NCCL all-gather")
        self.set_partition(comm if comm is not None else TorchDistComm(group))

    # -------------------------------------------------------------------------------------- construction
    def _validate_configuration(self, node_features, hidden_dims, num_diffusion_steps, attention_heads, dropout, graph_layers,
                                diffusion_schedule, activation, normalization, pooling, num_classes, regression_targets):
        validate_integer(node_features, 1, 10000)
        if not isinstance(hidden_dims, list) or len(hidden_dims) == 0:
            raise ValidationError("hidden_dims must be a non-empty list")
        for d in hidden_dims:
            validate_integer(d, 1, 10000)
        for i in range(1, len(hidden_dims)):
            if hidden_dims[i] > hidden_dims[i - 1]:
                warnings.warn(f"Hidden dimension {i} ({hidden_dims[i]}) > previous ({hidden_dims[i-1]})")
        validate_integer(num_diffusion_steps, 1, 1000)
        validate_enum(diffusion_schedule, ["linear", "cosine", "sigmoid"])
        validate_integer(attention_heads, 1, 32)
        if hidden_dims[-1] % attention_heads != 0:
            raise ValidationError(f"Hidden dim {hidden_dims[-1]} not divisible by attention heads {attention_heads}")
        validate_numeric(dropout, 0.0, 0.9)
        validate_integer(graph_layers, 1, 20)
        validate_enum(activation, ["relu", "gelu", "elu", "swish"])
        validate_enum(normalization, ["layer", "batch", "instance", "graph"])
        validate_enum(pooling, ["mean", "max", "attention", "set2set", "sort"])
        if num_classes is not None:
            validate_integer(num_classes, 2, 1000)
        validate_integer(regression_targets, 0, 100)
        for d in hidden_dims:
            if d % attention_heads != 0:
                raise ValidationError(f"Hidden dim {d} not divisible by attention heads {attention_heads}")
        if num_classes is None and regression_targets == 0:
            warnings.warn("No classification or regression targets specified - model will only do pretraining")

    def _create_pooling_layer(self, pooling, hidden_dim, attention_heads):
        if pooling == "mean":
            return GlobalMeanPool()
        if pooling == "max":
            return GlobalMaxPool()
        if pooling == "attention":
            return GlobalAttentionPool(hidden_dim, attention_heads)
        if pooling == "set2set":
            return GlobalSet2SetPool(hidden_dim)
        raise ValueError(f"Unknown pooling method: {pooling}")

    def _init_weights(self, module):
        if isinstance(module, nn.Linear):
            nn.init.xavier_uniform_(module.weight)
            if module.bias is not None:
                nn.init.constant_(module.bias, 0)
        elif isinstance(module, (nn.BatchNorm2d, nn.LayerNorm, nn.GroupNorm)):
            nn.init.constant_(module.weight, 1)
            nn.init.constant_(module.bias, 0)

    # -------------------------------------------------------------------------------------- validation
    def _validate_forward_inputs(self, data, mode, return_attention, return_embeddings, deferred=False):
        """deferred=True: the host-side structure checks run now, the validation kernel is enqueued without its read-back and its
        32-byte device buffer is returned — the caller reads it back behind the forward (`_check_validation_result`)."""
        if not isinstance(data, GRAPH_TYPES):
            raise ValidationError(f"Expected Data or Batch, got {type(data)}")
        for attr in ("x", "edge_index"):
            if getattr(data, attr, None) is None:
                raise ValidationError(f"Graph data missing required attribute: {attr}")
        validate_enum(mode, ["inference", "pretrain", "finetune"])
        validate_boolean(return_attention)
        validate_boolean(return_embeddings)
        x, ei = data.x, data.edge_index
        if x.dim() != 2:
            raise ValidationError(f"Node features must be 2D, got shape {x.shape}")
        if ei.dim() != 2 or ei.size(0) != 2:
            raise ValidationError(f"Edge index must be 2xN, got shape {ei.shape}")
        if x.size(1) != self.node_features:
            raise ValidationError(f"Expected {self.node_features} node features, got {x.size(1)}")
        if x.numel() == 0:
            raise ValidationError("Input features are empty")
        _require_cuda(x)
        if x.dtype != torch.float32 or ei.dtype != torch.int64 or not ei.is_cuda:
            raise ValidationError("x must be float32 and edge_index int64, both on the CUDA device")
        if deferred:
            return ops.validate_inputs_device(x.contiguous(), ei.contiguous(), self._num_nodes(data))
        flags, stats = ops.validate_inputs(x.contiguous(), ei.contiguous(), self._num_nodes(data))   # one kernel, one readback
        self._check_validation_result(flags, stats, x.numel())
        return None

    @staticmethod
    def _check_validation_result(flags, stats, n):
        """Host side of `_validate_input_features` / the edge-index range checks (models/dgdm_model.py:646-710) on the 32 bytes the
        validation kernel produced."""
        if flags[0]:
            raise ValidationError("Node features contain NaN values")
        if flags[1]:
            raise ValidationError("Node features contain infinity values")
        if flags[2]:
            raise ValidationError("Edge index contains invalid node indices")
        if flags[3]:
            raise ValidationError("Edge index contains negative node indices")
        amax, s, ss = stats[0], stats[1], stats[2]
        if amax > 1000:
            warnings.warn(f"Large feature values detected: max={amax}")
        if n > 1:
            var = max(ss - s * s / n, 0.0) / (n - 1)
            if var ** 0.5 < 1e-6:
                warnings.warn("Input features have very low variance")
        if n * 4 > 1e9:
            warnings.warn(f"Large feature tensor detected: {n * 4 / 1e9:.1f}GB")

    # -------------------------------------------------------------------------------------- forward
    def forward(self, data, mode="inference", return_attention=False, return_embeddings=False, *, rng=None, validate=True):
        """rng (pretrain only): dict(timesteps LongTensor [B], noise [N,D], target [N,D]) to inject the random draws
        (SURVEY §8c "RNG"); drawn from torch's generator when absent. validate=False skips the input checks."""
        try:
            if validate:
                self._validate_forward_inputs(data, mode, return_attention, return_embeddings)
        except Exception as e:
            raise ModelInferenceError(f"Input validation failed: {e}")
        if torch.is_grad_enabled() and mode in ("pretrain", "finetune") and any(p.requires_grad for p in self.parameters()):
            # training entry points (models/dgdm_model.py:405-506 under autograd): the differentiable path of train.py — hand-written
            # forward AND backward kernels for the GEMMs, the message passing and the spatial attention
            if self._comm is None and (getattr(data, "batch", None) is None or get_context(data, self._num_nodes(data)).sorted):
                from .train import train_forward
                try:
                    return train_forward(self, data, mode, return_embeddings, rng)
                except Exception as e:
                    raise ModelInferenceError(f"Forward pass failed: {e}") from e
            if not DGDMModel._warned_no_grad:
                DGDMModel._warned_no_grad = True
                warnings.warn("node-partitioned / unsorted-batch forwards run the inference kernels only: outputs carry no autograd graph")
        try:
            with torch.no_grad():
                return self._forward_impl(data, mode, return_attention, return_embeddings, rng)
        except ModelInferenceError:
            raise
        except Exception as e:
            raise ModelInferenceError(f"Forward pass failed: {e}") from e

    def graphed(self, data, mode="inference", return_embeddings=False, warmup=2, rebuild_structure=False, validate=False):
        """CUDA-graph replay of `forward(data, mode, return_embeddings=...)` for small graphs, where the ~150 kernel launches and
        ~100 ctypes calls of a forward cost more than the kernels (host-bound below ~10k nodes). The forward is validated and
        warmed up eagerly (weight images, CSR, kernel attributes), captured once, and the returned callable replays it:
        `run = model.graphed(d); out = run()` — `out` tensors are static buffers overwritten by the next replay; the graph reads
        `data.x` / `data.pos` / `data.edge_attr` in place, so new features may be copied into those tensors between replays
        (same graph structure). Single-GPU inference / finetune forwards (capturing the node-partitioned forward — NCCL all-gathers and
        symmetric-memory barriers inside the graph — was tried and dead-locked on a 2-GPU box: not supported).
        rebuild_structure=True keeps the CSR build inside the captured forward. validate=True captures the input-validation kernel
        as well: `run.check()` reads its 32 bytes back after a replay and raises exactly what `forward` raises for NaN / Inf features or
        out-of-range edge indices (deferred validation: one host synchronisation per forward instead of one in front of it)."""
        if self._comm is not None or mode == "pretrain":
            raise ModelInferenceError("graphed() supports single-GPU inference / finetune forwards")
        for _ in range(max(1, warmup)):
            self.forward(data, mode, False, return_embeddings)
        versions = self._weights_version()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = ops.LAUNCHES["kernels"]
        with torch.no_grad(), torch.cuda.graph(graph, capture_error_mode="thread_local"):
            if rebuild_structure and hasattr(data, "__dict__") and getattr(data, "batch", None) is None:
                data.__dict__.pop("_dgdm_ctx", None)          # (a batch vector needs a host read for the segment sizes: kept cached)
            vbuf = None
            if validate:
                vbuf = ops.validate_inputs_device(data.x.contiguous(), data.edge_index.contiguous(), self._num_nodes(data))
            static_out = self.forward(data, mode, False, return_embeddings, validate=False)
        model = self
        n_x = int(data.x.numel())

        def check():
            if vbuf is None:
                return
            host = vbuf.cpu()
            try:
                DGDMModel._check_validation_result(host[:4].tolist(), host[4:].view(torch.float32).tolist(), n_x)
            except Exception as e:
                raise ModelInferenceError(f"Input validation failed: {e}")

        def run():
            if model._weights_version() != versions:
                raise ModelInferenceError("parameters changed since graphed() captured the forward: capture again")
            graph.replay()
            return static_out
        run.graph, run.outputs, run.launches, run.check = graph, static_out, ops.LAUNCHES["kernels"] - l0, check
        run.validation_buffer, run.n_x = vbuf, n_x        # for callers that read the 32 bytes back on their own stream (GraphServer)
        return run

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _num_nodes(self, data):
        """N of the whole batch: x may hold only the owned rows of a node-partitioned graph."""
        if self._comm is None:
            return int(data.x.size(0))
        n = data.__dict__.get("_num_nodes") if hasattr(data, "__dict__") else None
        if n is None:
            pos, batch = getattr(data, "pos", None), getattr(data, "batch", None)
            n = pos.size(0) if pos is not None else (batch.numel() if batch is not None else data.x.size(0))
        return int(n)

    def _forward_impl(self, data, mode, return_attention, return_embeddings, rng):
        x = data.x.contiguous()
        N = self._num_nodes(data)
        ctx = get_context(data, N)
        batch = getattr(data, "batch", None)
        pos = getattr(data, "pos", None)
        edge_attr = getattr(data, "edge_attr", None)
        # ---- stage A inputs. The kernels index node-sorted segments; the reference's boolean masks (`x[batch == i]`,
        # models/dgdm_model.py:343-350) accept any node order, so an unsorted batch runs on a stably sorted copy.
        dA_x, dA_pos, ctxA, perm, inv = x, pos, ctx, None, None
        if batch is not None and not ctx.sorted:
            if self._comm is not None:
                raise ModelInferenceError("node-partitioned execution needs a batch vector sorted by graph id")
            perm = torch.sort(batch, stable=True).indices
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(perm.numel(), device=perm.device)
            srt = Data(x=x[perm].contiguous(), edge_index=inv[data.edge_index].contiguous(), pos=None if pos is None else pos[perm].contiguous())
            srt.batch = batch[perm].contiguous()
            dA_x, dA_pos, ctxA = srt.x, srt.pos, get_context(srt, N)
        part = None
        if self._comm is not None:
            from .parallel import Part, even_bounds
            if mode == "pretrain":
                raise ModelInferenceError("mode='pretrain' is not available for node-partitioned graphs (inference / finetune are)")
            part = Part(self._comm, even_bounds(N, self._comm.world))
            if dA_x.size(0) == N and part.n_local != N:
                dA_x = dA_x[part.a:part.b]
            elif dA_x.size(0) != part.n_local:
                raise ModelInferenceError(f"rank {self._comm.rank} owns rows [{part.a}, {part.b}) but data.x has {dA_x.size(0)} rows")
        h = self.feature_encoder(dA_x)
        h = self.graph_encoder(h, ctx=ctxA, edge_attr=edge_attr, part=part)["embeddings"]
        outputs, attn = {}, None
        if self.spatial_attention is not None and pos is not None:
            h, attn = self.spatial_attention.forward_nodes(h, dA_pos, ctxA.seg_ptr, ctxA.num_graphs, ctxA.seg_host, return_attention, part=part)
            if attn is not None and batch is None:
                attn = attn[0]
            # unsorted batch: the reference concatenates the per-graph attention outputs in GRAPH order (:358) and carries on with the
            # original edge_index / batch, i.e. row r now holds the r-th node of the sorted order. Reproduced as is: `h` stays sorted.
        elif perm is not None:
            h = h[inv]
        if self.hierarchical_processor is not None:
            h = self.hierarchical_processor(h, ctx=ctx, edge_attr=edge_attr, part=part)
        h_seg = h if perm is None else h[perm].contiguous()          # `node_embeddings[data.batch == i]` of the reference, graph by graph
        if mode == "pretrain":
            outputs.update(self._compute_diffusion_loss(h_seg, ctxA, rng))
        if isinstance(self.global_pool, GlobalAttentionPool):
            g = self.global_pool(h_seg, ctx=ctxA, part=part)
        else:
            h_pool = h_seg if part is None else part.gather(h_seg)
            g = self.global_pool(h_pool, batch, ctx=ctxA if batch is not None else None)
        if self.classification_head is not None and mode in ("inference", "finetune"):
            logits = self.classification_head(g)
            outputs["classification_logits"] = logits
            outputs["classification_probs"] = F.softmax(logits, dim=-1)
        if self.regression_head is not None and mode in ("inference", "finetune"):
            outputs["regression_outputs"] = self.regression_head(g)
        outputs["graph_embedding"] = g
        if return_embeddings:
            outputs["node_embeddings"] = h
            if part is not None:
                outputs["node_range"] = part.rows
        if return_attention and attn is not None:
            outputs["attention_weights"] = attn
        return outputs

    def _compute_diffusion_loss(self, h, ctx, rng):
        """models/dgdm_model.py:405-445 with P6: one timestep per graph, noise prediction MSE against an INDEPENDENT
        target draw (D13, reproduced)."""
        B, dev = ctx.num_graphs, h.device
        if rng is None:
            rng = {}
        t = rng.get("timesteps")
        if t is None:
            t = torch.randint(0, self.num_diffusion_steps, (B,), device=dev)
        noise = rng.get("noise")
        target = rng.get("target")
        noise = torch.randn_like(h) if noise is None else noise.to(dev)
        target = torch.randn_like(h) if target is None else target.to(dev)
        t_nodes = torch.repeat_interleave(t.to(dev), torch.tensor([ctx.seg_host[i + 1] - ctx.seg_host[i] for i in range(B)], device=dev))
        xn, eps = self.diffusion_layer(h, t_nodes, noise)
        losses = [F.mse_loss(eps[ctx.seg_host[i]:ctx.seg_host[i + 1]], target[ctx.seg_host[i]:ctx.seg_host[i + 1]]) for i in range(B)]
        a, b = ctx.seg_host[B - 1], ctx.seg_host[B]
        return {"diffusion_loss": torch.stack(losses).mean(), "noisy_embeddings": xn[a:b].unsqueeze(0)}

    def pretrain_step(self, data, mask_ratio=0.15, *, rng=None):
        """models/dgdm_model.py:447-506: entity masking + forward(mode='pretrain'); loss = diffusion loss (D12)."""
        masked = self._apply_entity_masking(data, mask_ratio, rng)
        outputs = self.forward(masked, mode="pretrain", rng=rng)
        outputs["total_pretrain_loss"] = outputs["diffusion_loss"]
        return outputs

    def _apply_entity_masking(self, data, mask_ratio, rng=None):
        masked = data.clone()
        n = data.x.size(0)
        k = int(n * mask_ratio)
        if k > 0:
            rng = rng or {}
            idx = rng.get("mask_indices")
            idx = torch.randperm(n)[:k] if idx is None else idx
            tok = rng.get("mask_token")
            tok = torch.randn(data.x.size(1)) if tok is None else tok
            node_mask = torch.zeros(n, dtype=torch.bool, device=data.x.device)
            node_mask[idx.to(data.x.device)] = True
            masked.x = data.x.clone()
            masked.x[node_mask] = tok.to(data.x.device)                      # D19 fixed: token moved to the data's device
            masked.node_mask = node_mask
        return masked

    def generate_embeddings(self, data, layer="final"):
        with torch.no_grad():
            out = self.forward(data, mode="inference", return_embeddings=True)
            if layer == "final":
                return out["graph_embedding"]
            if layer == "node":
                return out["node_embeddings"]
            raise ValueError(f"Unknown layer: {layer}")
