"""TEST INFRASTRUCTURE — loads the *real* reference (read-only tree at /root/reference) in this
container so that the oracle restatement (oracle/dgdm_oracle.py) can be pinned against it and golden
vectors can be generated (oracle/make_golden.py).  It cannot travel to the GPU box; nothing in the
product, `-m gpu` tests, smoke() or bench.py may import this module.

The reference forward does not execute as shipped.  `load(patches=...)` applies the
enumerated, individually toggleable patch set of SURVEY.md §8c as monkey-patches over the verbatim
reference files:

  P1  models/encoders.py:173-177     clamp the `dims[i]` index (DGDM-Large / 2-dim test models)
  P2  core/graph_layers.py:76-92     pad edge_attr with N zero rows when self-loops are appended
  P3  models/encoders.py:181-199     `transitions[i] = Linear(in,out)` between DynamicGraphLayer (returns
                                     node_dim, tests/test_basic.py:102) and LayerNorm(out_dim)
  P4  models/dgdm_model.py:339-342   hasattr(data,k) -> getattr(data,k,None) is not None
  P5  core/graph_layers.py:371-394   GraphUNet builds DynamicGraphLayer(node_dim=h, edge_dim=32, hidden_dim=h);
                                     edge_attr None -> zeros(E,32)
  P6  core/diffusion.py:94-104       run the diffusion layer on 2-D [N_i, D] (GroupNorm on 3-D fails)
  P7  builder/model edge_attr width  zero-pad edge features (2 for the builder, 32 for the model)

`monitor_operation` (utils/monitoring.py:479, gc.collect() per call) is left as shipped unless
stub_monitor=True.
"""
import contextlib
import importlib
import os
import sys

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DGDM_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "dgdm_histopath")):
    REF_ROOT = os.path.join(_HERE, "_ref")          # the copy staged by oracle/stage_ref.py (travels to the GPU box)
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
ALL_PATCHES = ("P1", "P2", "P3", "P4", "P5", "P6", "P7")

_state = {}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "dgdm_histopath"))


def is_live_tree():
    """True when REF_ROOT is the read-only reference checkout itself (build container), False for the staged copy."""
    return available() and os.path.abspath(REF_ROOT) != os.path.abspath(os.path.join(_HERE, "_ref"))


def set_monitor_stub(stub):
    """Switch the reference's `monitor_operation` (utils/monitoring.py:479: psutil sampling + gc.collect() around every stage,
    about half of its forward time at 2k nodes) between as-shipped (False) and a no-op (True) after load()."""
    mdl, orig, noop = _state["monitor"]
    mdl.monitor_operation = noop if stub else orig


def _pad_cols(t, width):
    if t is None or t.size(1) >= width:
        return t
    return torch.cat([t, t.new_zeros(t.size(0), width - t.size(1))], dim=1)


def load(patches=ALL_PATCHES, stub_monitor=True):
    """Import the reference package (once per process) and apply `patches`. Returns a namespace dict."""
    key = (tuple(sorted(patches)), stub_monitor)
    if _state.get("key") == key:
        return _state["ns"]
    if "key" in _state:
        raise RuntimeError("reference already loaded with a different patch set in this process")
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import warnings
    for p in (REF_ROOT, SHIM_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gl = importlib.import_module("dgdm_histopath.core.graph_layers")
        att = importlib.import_module("dgdm_histopath.core.attention")
        dif = importlib.import_module("dgdm_histopath.core.diffusion")
        enc = importlib.import_module("dgdm_histopath.models.encoders")
        mdl = importlib.import_module("dgdm_histopath.models.dgdm_model")
        dec = importlib.import_module("dgdm_histopath.models.decoders")
        tgb = importlib.import_module("dgdm_histopath.preprocessing.tissue_graph_builder")
        sp = importlib.import_module("dgdm_histopath.preprocessing.slide_processor")

    @contextlib.contextmanager
    def _noop(*a, **k):
        yield
    _state["monitor"] = (mdl, mdl.monitor_operation, _noop)
    if stub_monitor:
        mdl.monitor_operation = _noop

    if "P2" in patches:
        _conv_forward = gl.GraphConvolution.forward

        def conv_forward(self, x, edge_index, edge_attr=None, size=None):
            if edge_attr is not None and self.normalize and self.add_self_loops:
                edge_attr = torch.cat([edge_attr, edge_attr.new_zeros(x.size(0), edge_attr.size(1))], dim=0)
            return _conv_forward(self, x, edge_index, edge_attr, size)
        gl.GraphConvolution.forward = conv_forward

    if "P7" in patches:
        _dyn_forward = gl.DynamicGraphLayer.forward

        def dyn_forward(self, x, edge_index, edge_attr):
            return _dyn_forward(self, x, edge_index, _pad_cols(edge_attr, self.edge_dim))
        gl.DynamicGraphLayer.forward = dyn_forward

    if "P1" in patches or "P3" in patches:
        _enc_init = enc.GraphEncoder.__init__

        def enc_init(self, input_dim, hidden_dims, num_layers=4, **kw):
            dims = [input_dim] + list(hidden_dims)
            n_ok = min(num_layers, len(dims)) if "P1" in patches else num_layers
            _enc_init(self, input_dim, hidden_dims, num_layers=n_ok, **kw)
            heads, drop = kw.get("attention_heads", 8), kw.get("dropout", 0.1)
            for i in range(n_ok, num_layers):          # P1: layers past the end of `dims` are last->last
                self.graph_layers.append(gl.DynamicGraphLayer(
                    node_dim=dims[-1], edge_dim=32, hidden_dim=dims[-1], num_heads=heads, dropout=drop))
                self.norm_layers.append(self._get_norm_layer(kw.get("normalization", "layer"), dims[-1]))
            self.num_layers = num_layers
            if "P3" in patches:
                tr = []
                for i in range(num_layers):
                    a = dims[min(i, len(dims) - 1)]
                    b = dims[min(i + 1, len(dims) - 1)]
                    tr.append(nn.Linear(a, b) if a != b else nn.Identity())
                self.transitions = nn.ModuleList(tr)

        def enc_forward(self, x, edge_index, edge_attr=None, batch=None):
            h, outs = x, []
            for i, (layer, norm) in enumerate(zip(self.graph_layers, self.norm_layers)):
                ea = edge_attr if (edge_attr is not None and self.use_edge_features) else \
                    torch.zeros(edge_index.size(1), 32, device=h.device)
                h = layer(h, edge_index, ea)
                if hasattr(self, "transitions"):
                    h = self.transitions[i](h)
                h = self.dropout(self.activation(norm(h)))
                outs.append(h.clone())
            return {"embeddings": self.output_proj(h), "layer_outputs": outs, "num_nodes": x.size(0)}
        enc.GraphEncoder.__init__ = enc_init
        enc.GraphEncoder.forward = enc_forward

    if "P5" in patches:
        _unet_forward = gl.GraphUNet.forward

        def unet_init(self, in_channels, hidden_channels, out_channels, depth=3, pool_ratios=None,
                      sum_res=True, act="relu"):
            nn.Module.__init__(self)
            self.in_channels, self.hidden_channels, self.out_channels = in_channels, hidden_channels, out_channels
            self.depth, self.sum_res = depth, sum_res
            ratios = pool_ratios or [0.5] * depth
            F = torch.nn.functional
            self.act = {"relu": F.relu, "gelu": F.gelu}.get(act, F.elu)
            mk = lambda nd: gl.DynamicGraphLayer(node_dim=nd, edge_dim=32, hidden_dim=hidden_channels)
            self.down_convs = nn.ModuleList([mk(in_channels)] + [mk(hidden_channels) for _ in range(depth)])
            self.pools = nn.ModuleList([gl.AdaptiveGraphPooling(hidden_channels, ratio=r) for r in ratios])
            self.bottom_conv = mk(hidden_channels)
            up_dim = hidden_channels if sum_res else 2 * hidden_channels
            self.up_convs = nn.ModuleList([mk(up_dim) for _ in range(depth)])
            self.final_conv = nn.Linear(hidden_channels, out_channels)

        def unet_forward(self, x, edge_index, edge_attr=None, batch=None):
            if edge_attr is None:
                edge_attr = x.new_zeros(edge_index.size(1), 32)
            return _unet_forward(self, x, edge_index, _pad_cols(edge_attr, 32), batch)
        gl.GraphUNet.__init__ = unet_init
        gl.GraphUNet.forward = unet_forward

    if "P4" in patches:
        mdl.hasattr = lambda o, k: getattr(o, k, None) is not None   # shadows the builtin inside dgdm_model only

    if "P6" in patches:
        _dif_forward = dif.DiffusionLayer.forward

        def dif_forward(self, x_start, timesteps=None, noise=None, condition=None):
            if x_start.dim() == 3 and x_start.size(0) == 1 and timesteps is not None:
                x2 = x_start[0]
                t = timesteps.reshape(-1)[:1].repeat(x2.size(0))
                n2 = noise[0] if noise is not None else None
                xn, pn = _dif_forward(self, x2, t, n2, condition)
                return xn.unsqueeze(0), pn.unsqueeze(0)
            return _dif_forward(self, x_start, timesteps, noise, condition)
        dif.DiffusionLayer.forward = dif_forward

    if "P7" in patches:
        _to_pyg = tgb.TissueGraphBuilder._to_pytorch_geometric

        def to_pyg(self, nodes, edges, slide_data):
            import numpy as np
            for e in edges:
                if e.features is not None and len(e.features) < 2:
                    e.features = np.concatenate([e.features, np.zeros(2 - len(e.features))])
            return _to_pyg(self, nodes, edges, slide_data)
        tgb.TissueGraphBuilder._to_pytorch_geometric = to_pyg

    import torch_geometric.data as tgd
    ns = dict(gl=gl, att=att, dif=dif, enc=enc, mdl=mdl, dec=dec, tgb=tgb, sp=sp,
              Data=tgd.Data, Batch=tgd.Batch, DGDMModel=mdl.DGDMModel,
              TissueGraphBuilder=tgb.TissueGraphBuilder, patches=tuple(patches))
    _state["key"], _state["ns"] = key, ns
    return ns
