"""Size-independent properties at the BASELINE.json sizes (100k-node graph, 4 x 50k batch), where the CPU oracle cannot run
(the reference's dense attention would need 320 GB): determinism, softmax normalisation, linearity, permutation
equivariance of message passing, batch == per-graph for graph-local stages."""
import pytest
import torch

from dgdm_histopath_lab_b200 import Batch, Data, ops, synth
from dgdm_histopath_lab_b200.graph import GraphContext
from util import BASE_CFG, make_model, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def base_model():
    return make_model(BASE_CFG).to(DEV)


@pytest.fixture(scope="module")
def graph100k():
    return synth.tile_graph(100000, 768, seed=1234).to(DEV)


def test_forward_100k_is_deterministic_bitwise_and_finite(base_model, graph100k):
    """no atomics on the data path, fixed accumulation order: two runs agree bit for bit (our race detector)"""
    a = base_model(graph100k, return_embeddings=True)
    b = base_model(graph100k, return_embeddings=True)
    assert torch.equal(a["node_embeddings"], b["node_embeddings"]) and torch.equal(a["graph_embedding"], b["graph_embedding"])
    assert torch.isfinite(a["node_embeddings"]).all() and a["node_embeddings"].shape == (100000, 128)


def test_attention_100k_rows_are_convex_combinations():
    """softmax rows sum to 1: with v = 1 the output must be 1. This is the worst case for the tensor core's truncating fp32
    accumulation (all terms positive): an uncapped 781-block chain measured -3.5e-4; with the chain capped at 48 key blocks
    per CTA (attention_tc.cu: TC_MAX_CHAIN) the bias stays below 3e-5, 3x inside the 1e-4 parity bar."""
    n, D, H = 100000, 128, 8
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(n, 3 * D, generator=g).to(DEV)
    qkv[:, 2 * D:] = 1.0
    pos = torch.rand(n, 2, generator=g).to(DEV)
    seg = torch.tensor([0, n], dtype=torch.int32, device=DEV)
    out = ops.spatial_attention(qkv, pos, seg, 1, H, 1.0)
    assert (out - 1.0).abs().max().item() < 3e-5


def test_spmm_100k_linearity_and_permutation_equivariance():
    d = synth.tile_graph(100000, 8, seed=5)
    ei = d.edge_index.to(DEV)
    n, C = 100000, 128
    ctx = GraphContext(ei, None, n)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(n, C, generator=g).to(DEV), torch.randn(n, C, generator=g).to(DEV)
    ax, ay, axy = ops.spmm_gcn(ctx.csr, x), ops.spmm_gcn(ctx.csr, y), ops.spmm_gcn(ctx.csr, 2.0 * x + y)
    assert rel_err(axy, 2.0 * ax.double() + ay.double())[0] < 2e-6
    # relabel the nodes: P A P^T (P x) = P (A x)
    perm = torch.randperm(n, generator=g).to(DEV)
    inv = torch.empty_like(perm); inv[perm] = torch.arange(n, device=DEV)
    ctx_p = GraphContext(inv[ei], None, n)                       # node i becomes node inv[i]
    xp = torch.empty_like(x); xp[inv] = x
    out_p = ops.spmm_gcn(ctx_p.csr, xp)
    assert rel_err(out_p[inv], ax)[0] < 2e-6                      # same neighbour multiset per node; only the summation order differs


def test_batch_4x50k_encoder_and_attention_equal_per_graph(base_model):
    """BASELINE.json configs[1]: the encoder and the spatial attention are graph-local, so a 4-graph batch must reproduce the
    four single-graph results (the batch-global top-k of the reference's pooling, D16, only enters in the U-Net)."""
    graphs = [synth.tile_graph(50000, 768, seed=100 + i) for i in range(4)]
    b = Batch.from_data_list(graphs).to(DEV)
    m = base_model
    from dgdm_histopath_lab_b200.graph import get_context
    ctx = get_context(b)
    h = m.graph_encoder(m.feature_encoder(b.x), ctx=ctx)["embeddings"]
    hb, _ = m.spatial_attention.forward_nodes(h, b.pos, ctx.seg_ptr, ctx.num_graphs, ctx.seg_host)
    for i in (0, 3):
        d = graphs[i].to(DEV)
        c1 = get_context(d)
        h1 = m.graph_encoder(m.feature_encoder(d.x), ctx=c1)["embeddings"]
        a1, _ = m.spatial_attention.forward_nodes(h1, d.pos, c1.seg_ptr, 1, c1.seg_host)
        sl = slice(50000 * i, 50000 * (i + 1))
        assert torch.equal(h[sl], h1)                             # message passing / GEMMs: order-identical
        assert rel_err(hb[sl], a1)[0] < 3e-5                      # attention: the key-split count (accumulation chains) differs with the batch
    out = m(b)
    assert out["graph_embedding"].shape == (4, 128) and torch.isfinite(out["graph_embedding"]).all()


def test_attention_100k_sampled_rows_vs_float64_dense():
    """Full-size parity on SAMPLED query rows: the reference cannot form the 100k x 100k scores, but 96 rows of them fit anywhere —
    softmax(q k^T / sqrt(hd) - |dp| / T) v in float64 (torch, on the device, as the checker) vs the tcgen05 kernel's rows.
    Bar: 5e-5 max-scaled (the 1e-4 parity bar with margin; includes the capped tensor-core accumulation bias)."""
    n, D, H, hd = 100000, 128, 8, 16
    g = torch.Generator().manual_seed(17)
    qkv = torch.randn(n, 3 * D, generator=g).to(DEV)
    pos = torch.rand(n, 2, generator=g).to(DEV)
    seg = torch.tensor([0, n], dtype=torch.int32, device=DEV)
    out = ops.spatial_attention(qkv, pos, seg, 1, H, 0.8)
    rows = torch.randint(0, n, (96,), generator=g).to(DEV)
    q = qkv[rows, :D].double().view(-1, H, hd)
    k = qkv[:, D:2 * D].double().view(n, H, hd)
    v = qkv[:, 2 * D:].double().view(n, H, hd)
    s = torch.einsum("rhd,nhd->rhn", q, k) / hd ** 0.5 - (torch.cdist(pos[rows].double(), pos.double()) / 0.8)[:, None, :]
    ref = torch.einsum("rhn,nhd->rhd", torch.softmax(s, dim=-1), v).reshape(-1, D)
    assert rel_err(out[rows], ref)[0] < 5e-5


def test_encoder_layer_100k_sampled_rows_vs_float64():
    """One DynamicGraphLayer (two GraphConvolutions + output projection + residual + LayerNorm) at 100k nodes / 1.6M edges:
    sampled destination rows recomputed in float64 from the layer's own weights."""
    from dgdm_histopath_lab_b200.core import DynamicGraphLayer
    import torch.nn.functional as F
    torch.manual_seed(3)
    n, C = 100000, 128
    d = synth.tile_graph(n, 8, seed=6)
    ei = d.edge_index.to(DEV)
    layer = DynamicGraphLayer(C, 32, C).to(DEV).eval()
    x = torch.randn(n, C, generator=torch.Generator().manual_seed(4)).to(DEV)
    out = layer(x, ei)
    # float64 restatement on the device (dense ops over all nodes are cheap; only the comparison is sampled)
    deg = torch.bincount(ei[1], minlength=n).double() + 1.0
    dinv = deg.pow(-0.5)

    def conv(conv_mod, h):
        hl = h @ conv_mod.node_lin.weight.double().t()
        msg = hl[ei[0]] * (dinv[ei[0]] * dinv[ei[1]])[:, None]
        agg = torch.zeros(n, hl.size(1), dtype=torch.float64, device=DEV).index_add_(0, ei[1], msg) + hl * (dinv * dinv)[:, None]
        return agg + conv_mod.bias.double()

    h1 = F.gelu(conv(layer.graph_conv1, x.double()))
    h2 = F.gelu(conv(layer.graph_conv2, h1))
    y = h2 @ layer.output_proj.weight.double().t() + layer.output_proj.bias.double() + x.double()
    ref = F.layer_norm(y, (C,), layer.norm1.weight.double(), layer.norm1.bias.double(), 1e-5)
    rows = torch.randint(0, n, (4096,), generator=torch.Generator().manual_seed(5)).to(DEV)
    assert rel_err(out[rows], ref[rows])[0] < 1e-5


def test_unet_topk_flip_report_100k():
    """SURVEY §7 hard part / VERDICT r1 #1d: at the BASELINE size the three discontinuous top-50 % selections (100k / 50k / 25k
    scores) of the product path are compared with a float64 restatement on the device. A flip is REPORTED (a node whose fp64
    score sits within fp32 noise of the k-th boundary legitimately swaps); the test fails only when the selections agree and
    the U-Net output still differs by more than the 1e-4 bar."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from flip_report import flip_report
    rep = flip_report(100000)
    print("flip report:", json.dumps(rep))
    for lv in rep["levels"]:
        if lv["perm_delta"]:
            assert lv["kth_boundary_gap_fp64"] < 1e-5, f"nodes flipped although the k-th boundary gap is {lv['kth_boundary_gap_fp64']}"
    if rep["flips_total"] == 0:
        assert rep["unet_out_rel_err_vs_fp64"] < 1e-4
