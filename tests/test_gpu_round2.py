"""GPU parity for the public surface the round-1 review found untested: edge attributes through the WHOLE model (GraphUNet
pooled levels), the serving pipeline build_graph -> forward, pretrain_step / entity masking / generate_embeddings, and the
public MultiHeadAttention / SpatialAttention calls (masks, key padding, cross attention, per-head weights, B > 1).
Every case is checked against the golden vectors recorded from the real reference (oracle/make_golden.py --round2) AND the
CPU oracle. Tolerance: rel 1e-4 (fp32 path), index outputs exact."""
import numpy as np
import pytest
import torch

from dgdm_histopath_lab_b200 import Batch, Data, synth
from dgdm_histopath_lab_b200 import PatchInfo, SlideData, TissueGraphBuilder
from oracle import dgdm_oracle as O
from util import BUILDER_MODEL_CFG, SMALL_CFG, golden, make_model, oracle_cfg, rel_err, sd_cpu

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


def _edge_attrs(b):
    g = torch.Generator().manual_seed(5)
    return torch.randn(b.edge_index.size(1), 32, generator=g), torch.rand(b.edge_index.size(1), 2, generator=g), g


def test_full_model_with_edge_attr_vs_golden():
    """core/graph_layers.py:404-453 threads edge_attr through every U-Net level; models/dgdm_model.py:371-374 always passes it."""
    g = golden("edge_attr_model.npz")
    m = make_model(SMALL_CFG).to(DEV)
    b = synth.tile_batch(2, 150, 48, seed=11)
    ea32, ea2, gen = _edge_attrs(b)
    for tag, ea in (("ea32", ea32), ("ea2", ea2)):
        bb = b.clone()
        bb.edge_attr = ea
        out = m(bb.to(DEV), return_embeddings=True)
        assert rel_err(out["node_embeddings"], g[tag + "_node_embeddings"])[0] < TOL, tag
        assert rel_err(out["graph_embedding"], g[tag + "_graph_embedding"])[0] < TOL, tag
        assert rel_err(out["classification_logits"], g[tag + "_logits"])[0] < TOL, tag
    d1 = synth.tile_graph(257, 32, seed=5)
    ea1 = torch.randn(d1.edge_index.size(1), 32, generator=gen)
    u, perms = m.hierarchical_processor(d1.x.to(DEV), d1.edge_index.to(DEV), ea1.to(DEV), return_perms=True)
    for i, pm in enumerate(perms):
        assert np.array_equal(pm.cpu().numpy(), g[f"unet1_perm{i}"]), f"top-k selection differs at level {i}"
    assert rel_err(u, g["unet1_out"])[0] < 2e-5


def test_edge_attr_is_not_ignored():
    m = make_model(SMALL_CFG).to(DEV)
    b = synth.tile_batch(2, 150, 48, seed=11)
    ea32, _, _ = _edge_attrs(b)
    base = m(b.to(DEV), return_embeddings=True)["node_embeddings"]
    bb = b.clone(); bb.edge_attr = ea32
    with_attr = m(bb.to(DEV), return_embeddings=True)["node_embeddings"]
    bz = b.clone(); bz.edge_attr = torch.zeros_like(ea32)
    zeros = m(bz.to(DEV), return_embeddings=True)["node_embeddings"]
    assert torch.equal(zeros, base), "zero attributes == no attributes (edge_lin has no bias)"
    assert (with_attr - base).abs().max() > 1e-3


def _slide_600():
    n = 600
    feats = synth.clustered_features(n, d=64, num_clusters=12, seed=21, sigma_lo=0.05, sigma_hi=0.11)
    xs, ys, dims = synth.jittered_grid_coords(n, seed=21, extent=3000)
    patches = [PatchInfo(int(xs[i]), int(ys[i]), 0, 20.0, f"s_patch_{i:06d}", 0.9, feats[i]) for i in range(n)]
    return SlideData("s", patches, {"dimensions": dims})


def test_build_graph_then_forward_vs_golden():
    """the reference's serving pipeline (evaluation/predictor.py:188-215): TissueGraphBuilder.build_graph -> DGDMModel.forward"""
    gf = golden("builder_forward_600.npz")
    gph = TissueGraphBuilder(feature_extractor="none").build_graph(_slide_600())
    assert gph.edge_attr.shape[1] == 2 and gph.edge_attr.shape[0] == gph.edge_index.shape[1]
    m = make_model(BUILDER_MODEL_CFG, seed=3).to(DEV)
    out = m(gph.to(DEV), return_embeddings=True)
    assert rel_err(out["node_embeddings"], gf["node_embeddings"])[0] < TOL
    assert rel_err(out["graph_embedding"], gf["graph_embedding"])[0] < TOL
    assert rel_err(out["classification_logits"], gf["logits"])[0] < TOL


def test_builder_output_without_edges_runs():
    """nothing passes the threshold -> the builder emits edge_index (2,0), edge_attr (0,1). The reference's own validation raises
    on it (edge_index.max() of an empty tensor, models/dgdm_model.py:676); here it runs and equals the oracle."""
    n = 40
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((n, 64)).astype(np.float32)                     # unrelated features: cosine << 0.7
    patches = [PatchInfo(int(i * 5000), int((i * 7919) % 90000), 0, 20.0, f"s_patch_{i:06d}", 0.9, feats[i]) for i in range(n)]
    gph = TissueGraphBuilder(feature_extractor="none").build_graph(SlideData("s", patches, {"dimensions": (200000, 100000)}))
    assert gph.edge_index.shape == (2, 0) and gph.edge_attr.shape == (0, 1)
    m = make_model(BUILDER_MODEL_CFG, seed=3)
    ref = O.dgdm_forward(sd_cpu(m), oracle_cfg(BUILDER_MODEL_CFG), gph.x, gph.edge_index, gph.edge_attr, None, gph.pos)
    out = m.to(DEV)(gph.to(DEV), return_embeddings=True)
    assert rel_err(out["node_embeddings"], ref["node_embeddings"])[0] < TOL


def test_pretrain_step_vs_golden_and_masking_replay():
    """models/dgdm_model.py:447-506 with the reference's RNG order replayed (randperm, mask token, timesteps, noise, target)."""
    g = golden("pretrain_step.npz")
    m = make_model(SMALL_CFG).to(DEV)
    b = synth.tile_batch(2, 150, 48, seed=11).to(DEV)
    rng = dict(mask_indices=torch.tensor(g["mask_indices"]), mask_token=torch.tensor(g["mask_token"]),
               timesteps=torch.tensor(g["timesteps"]), noise=torch.tensor(g["noise"]), target=torch.tensor(g["target"]))
    out = m.pretrain_step(b, mask_ratio=0.15, rng=rng)
    ref = float(g["total_pretrain_loss"])
    assert abs(float(out["total_pretrain_loss"]) - ref) < TOL * abs(ref)
    assert float(out["total_pretrain_loss"]) == float(out["diffusion_loss"])        # D12: reconstruction branch is dead
    assert rel_err(out["noisy_embeddings"], g["noisy_embeddings"])[0] < TOL
    assert "classification_logits" not in out
    # without injected draws the masking consumes torch's CPU generator in the reference's order (:490-497)
    torch.manual_seed(123)
    masked = m._apply_entity_masking(b, 0.15)
    idx = torch.tensor(g["mask_indices"])
    want = torch.zeros(300, dtype=torch.bool); want[idx] = True
    assert torch.equal(masked.node_mask.cpu(), want)
    assert torch.equal(masked.x[idx.to(DEV)].cpu(), torch.tensor(g["mask_token"]).expand(idx.numel(), -1))
    keep = ~want
    assert torch.equal(masked.x.cpu()[keep], b.x.cpu()[keep]) and not hasattr(b, "node_mask")
    # un-injected pretrain_step: finite loss, right shapes
    o2 = m.pretrain_step(b)
    assert torch.isfinite(o2["total_pretrain_loss"]) and o2["noisy_embeddings"].shape == (1, 150, 32)
    assert m._apply_entity_masking(b, 0.0).x is b.x or torch.equal(m._apply_entity_masking(b, 0.0).x, b.x)


def test_generate_embeddings():
    g = golden("small_model.npz")
    m = make_model(SMALL_CFG).to(DEV)
    b = synth.tile_batch(2, 150, 48, seed=11).to(DEV)
    assert rel_err(m.generate_embeddings(b, "final"), g["graph_embedding"])[0] < TOL
    assert rel_err(m.generate_embeddings(b, "node"), g["node_embeddings"])[0] < TOL
    with pytest.raises(ValueError):
        m.generate_embeddings(b, "bogus")


def test_public_multi_head_attention_vs_golden():
    """MultiHeadAttention.forward (core/attention.py:73-181): self / masked / cross attention, key padding, distinct value,
    head-averaged and per-head weights."""
    g = golden("mha_public.npz")
    m = make_model(SMALL_CFG).to(DEV)
    att = m.spatial_attention.attention
    q, k, v = (torch.tensor(g[n]).to(DEV) for n in ("q", "k", "v"))
    o, w = att(q)
    assert o.shape == (2, 37, 32) and w.shape == (2, 37, 37)
    assert rel_err(o, g["self_out"])[0] < 2e-5 and rel_err(w, g["self_w"])[0] < 2e-5
    o_flash, w_none = att(q, need_weights=False)                                    # flash path, all batch elements in one launch
    assert w_none is None and rel_err(o_flash, g["self_out"])[0] < 2e-5
    o, w = att(q, attn_mask=torch.tensor(g["attn_mask_sq"]).to(DEV), average_attn_weights=False)
    assert w.shape == (4, 37, 37)
    assert rel_err(o, g["selfmask_out"])[0] < 2e-5 and rel_err(w, g["selfmask_w"])[0] < 2e-5
    o, w = att(q, k, v, key_padding_mask=torch.tensor(g["key_padding_mask"]).to(DEV), attn_mask=torch.tensor(g["attn_mask"]).to(DEV))
    assert w.shape == (2, 37, 53)
    assert rel_err(o, g["cross_out"])[0] < 2e-5 and rel_err(w, g["cross_w"])[0] < 2e-5
    assert float(w[0, :, 40:].abs().max()) == 0.0 and float(w[1, :, :5].abs().max()) == 0.0
    o, w = att(q, k, need_weights=False)                                            # cross attention on the flash path (L_q != L_k)
    assert w is None and rel_err(o, g["cross_kv_out"])[0] < 2e-5
    # boolean mask == -inf additive mask; a fully masked row is NaN like torch's softmax
    bm = torch.zeros(37, 37, dtype=torch.bool, device=DEV); bm[:, 20:] = True
    fm = torch.zeros(37, 37, device=DEV).masked_fill(bm, float("-inf"))
    o1, w1 = att(q, attn_mask=bm)
    o2, w2 = att(q, attn_mask=fm)
    assert torch.equal(o1, o2) and torch.equal(w1, w2)
    ref, _ = O.multi_head_attention(sd_cpu(m), "spatial_attention.attention", q.cpu(), num_heads=2, attn_mask=bm.cpu())
    assert rel_err(o1, ref)[0] < 2e-5
    allm = torch.ones(37, 37, dtype=torch.bool, device=DEV)
    assert torch.isnan(att(q, attn_mask=allm)[1]).all()
    with pytest.raises(RuntimeError):
        att(q, attn_mask=torch.zeros(3, 37, 37, device=DEV))                        # not broadcastable against [2, 2, 37, 37]


def test_public_spatial_attention_batched_vs_golden():
    """SpatialAttention.forward with B > 1 (core/attention.py:285-327): global PE min/max, bias broadcast as the reference does."""
    g = golden("mha_public.npz")
    m = make_model(SMALL_CFG).to(DEV)
    x, pos = torch.tensor(g["spatial_x"]).to(DEV), torch.tensor(g["spatial_pos"]).to(DEV)
    o, w = m.spatial_attention(x, pos)
    assert o.shape == (2, 45, 32) and w.shape == (2, 45, 45)
    assert rel_err(o, g["spatial_out"])[0] < 2e-5 and rel_err(w, g["spatial_w"])[0] < 2e-5
    # B == 1 equals the per-graph node path the model uses
    o1, w1 = m.spatial_attention(x[:1], pos[:1])
    seg = torch.tensor([0, 45], dtype=torch.int32, device=DEV)
    on, wn = m.spatial_attention.forward_nodes(x[0].contiguous(), pos[0].contiguous(), seg, 1, [0, 45], True)
    assert rel_err(o1[0], on)[0] < 2e-5 and rel_err(w1[0], wn[0])[0] < 2e-5
    with pytest.raises(RuntimeError):
        m.spatial_attention(x.repeat(2, 1, 1)[:3], pos.repeat(2, 1, 1)[:3])         # B = 3, H = 2: the reference's broadcast fails


def test_weight_cache_follows_parameter_updates():
    """cached fp16 weight images are re-made when a parameter changes; q|k|v planes die with their model."""
    import gc
    from dgdm_histopath_lab_b200 import core
    m = make_model(SMALL_CFG).to(DEV)
    b = synth.tile_batch(2, 150, 48, seed=11).to(DEV)
    a0 = m(b, return_embeddings=True)["node_embeddings"].clone()
    with torch.no_grad():
        m.spatial_attention.attention.q_proj.weight.mul_(1.5)                      # bumps _version
    a1 = m(b, return_embeddings=True)["node_embeddings"].clone()
    assert (a1 - a0).abs().max() > 1e-4
    m.spatial_attention.attention.k_proj.weight.data.mul_(0.5)                      # .data write: invisible to _version
    core.invalidate_weight_cache()
    a2 = m(b, return_embeddings=True)["node_embeddings"]
    ref = O.dgdm_forward(sd_cpu(m), oracle_cfg(SMALL_CFG), b.x.cpu(), b.edge_index.cpu(), None, b.batch.cpu(), b.pos.cpu())
    assert rel_err(a2, ref["node_embeddings"])[0] < TOL
    n_before = len(core._CAT_CACHE._d)
    del m, a2
    gc.collect()
    assert len(core._CAT_CACHE._d) < n_before, "q|k|v planes must be released with the model"


@pytest.mark.parametrize("n", [10000, 25000])
def test_builder_full_edge_list_vs_sklearn_oracle(n):
    """BASELINE configs[2] exactness on checkable sizes: the WHOLE edge list (indices AND order) against oracle/builder_oracle.py
    (scikit-learn's NearestNeighbors / cosine_similarity exactly as the reference calls them, tissue_graph_builder.py:290-365),
    768-d clustered features whose within-cluster cosine straddles the 0.7 threshold, single pass and 3-shard row partition."""
    from oracle import builder_oracle
    feats = synth.clustered_features(n, d=768, num_clusters=max(8, n // 250), seed=31)
    # continuous jittered-grid coordinates: integer pixel coordinates produce EXACT 2-D distance ties (equal dx^2 + dy^2), whose
    # order in scikit-learn's kd-tree is implementation-defined (SURVEY 7 "hard parts"); SURVEY 8(d) config 3 asks for no ties
    rng = np.random.default_rng(31)
    w_ = int(np.ceil(np.sqrt(n)))
    ii = np.arange(n)
    coords = np.stack([((ii % w_) + 0.5 + rng.uniform(-0.15, 0.15, n)) / (w_ + 1), ((ii // w_) + 0.5 + rng.uniform(-0.15, 0.15, n)) / (w_ + 1)], axis=1)
    ei, ea, et = builder_oracle.build_edges(feats, coords)
    tb = TissueGraphBuilder(feature_extractor="none")
    f = torch.as_tensor(feats).to(DEV)
    c = torch.as_tensor(coords).to(DEV)
    e, stats = tb.edge_arrays_sharded(f, c, 0, 1)
    rows = -(-n // 3)
    shards = [tb.shard_candidates(f, c, q0, min(n, q0 + rows) - q0)[:2] for q0 in range(0, n, rows)]
    e3 = tb.merge_shards(shards, n)
    thr = 0.7
    # pairs whose cosine lies within 2e-6 of the threshold in EITHER direction: their keep/drop decision — and, when only one
    # direction survives, the first-seen POSITION of the edge — depends on the last bits of the cosine arithmetic
    from dgdm_histopath_lab_b200 import builder as B
    kidx, kcos, _ = B.knn_cosine(f, 17)
    nr, nc_ = torch.nonzero((kcos[:, 1:].double() - thr).abs() < 2e-6, as_tuple=True)
    near = {(min(int(i), int(kidx[i, j + 1])), max(int(i), int(kidx[i, j + 1]))) for i, j in zip(nr.tolist(), nc_.tolist())}
    ref_pairs = list(zip(ei[0, 0::2].tolist(), ei[1, 0::2].tolist()))
    ref_w = {pq: float(w) for pq, w in zip(ref_pairs, ea[0::2, 0])}
    for tag, ed in (("single", e), ("3 shards", e3)):
        src, tgt = ed["src"].cpu().numpy(), ed["tgt"].cpu().numpy()
        ours = list(zip(src.tolist(), tgt.tolist()))
        our_w = {pq: float(w) for pq, w in zip(ours, ed["f0"].cpu().numpy())}
        und = lambda pq: (min(pq), max(pq))
        so, sr = {und(p) for p in ours}, {und(p) for p in ref_pairs}
        diff = so ^ sr
        # the only admissible differences are pairs whose weight sits within 1e-6 of the threshold: scikit-learn's cosine is an fp32 sgemm
        # of normalised rows, ours an fp64-accumulated dot rounded once, so a pair at 0.7 +- 1e-7 can land on either side (north_star:
        # "pairs within 1e-6 of edge_threshold counted and reported")
        for pq in diff:
            w = [d_[k] for d_ in (our_w, ref_w) for k in (pq, pq[::-1]) if k in d_]
            assert w and all(abs(x - thr) < 2e-6 or abs(np.exp(-10 * x) - thr) < 2e-6 for x in w), (tag, pq, w)
        assert len(diff) <= stats["near_threshold_spatial"] + stats["near_threshold_morphological"], (tag, len(diff), stats)
        assert diff <= near, (tag, diff - near)
        keep_o = [p for p in ours if und(p) not in near]
        keep_r = [p for p in ref_pairs if und(p) not in near]
        if keep_o != keep_r:
            first = next((i for i, (a_, b_) in enumerate(zip(keep_o, keep_r)) if a_ != b_), min(len(keep_o), len(keep_r)))
            print(tag, "first difference at", first, keep_o[first - 2:first + 3], keep_r[first - 2:first + 3], len(keep_o), len(keep_r),
                  "only ours", list(set(keep_o) - set(keep_r))[:5], "only ref", list(set(keep_r) - set(keep_o))[:5])
        assert keep_o == keep_r, f"{tag}: edge list (indices and order) differs from the sklearn oracle beyond the near-threshold pairs"
        ty = dict(zip(ours, ed["type"].cpu().numpy().tolist()))
        assert all(ty[p] == t for p, t in zip(ref_pairs, et[0::2].tolist()) if und(p) not in near)
        assert all(abs(our_w[p] - ref_w[p]) < 2e-6 for p in keep_r)
    ndiff = len(diff)
    print(f"n={n}: {ei.shape[1] // 2} undirected edges, identical to the oracle up to {ndiff} near-threshold pair(s); counted: {stats}")
    assert stats["near_threshold_spatial"] >= 0 and stats["near_threshold_morphological"] >= 0


def test_contrastive_loss_vs_reference_formula():
    """ContrastiveLoss._batch_contrastive_loss (training/losses.py:123-176) — the dense float64 restatement of the reference
    formula vs the chunked tensor-core path that never forms the N x N matrix."""
    from dgdm_histopath_lab_b200 import ContrastiveLoss
    g = torch.Generator().manual_seed(4)
    n, d = 1337, 128
    emb = torch.randn(n, d, generator=g)
    batch = torch.sort(torch.randint(0, 5, (n,), generator=g)).values
    batch[-1] = 7                                                          # a graph with a single node: no positives, dropped (:168-170)

    def dense(e, b, temp, reduction):
        e = torch.nn.functional.normalize(e.double(), dim=1)
        s = e @ e.t() / temp
        pm = (b[None, :] == b[:, None]).double()
        pm.fill_diagonal_(0)
        ex = torch.exp(s)
        pos = torch.clamp((ex * pm).sum(1), min=1e-8)
        loss = -torch.log(pos / ex.sum(1))[pm.sum(1) > 0]
        return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss
    for temp, red in ((0.1, "mean"), (0.5, "sum"), (0.07, "none")):
        out = ContrastiveLoss(temperature=temp, reduction=red)(emb.to(DEV), batch.to(DEV))
        ref = dense(emb, batch, temp, red)
        assert rel_err(out.reshape(-1), ref.reshape(-1))[0] < 2e-5, (temp, red)
    one = ContrastiveLoss()(emb.to(DEV), None)                             # batch_indices None: every other node is a positive
    assert rel_err(one.reshape(-1), dense(emb, torch.zeros(n, dtype=torch.long), 0.1, "mean").reshape(-1))[0] < 2e-5
    with pytest.raises(Exception):
        ContrastiveLoss()(emb, batch)                                      # CPU tensors: no fallback


def test_graph_file_with_stored_csr_and_device_collate(tmp_path):
    """a .dgdm file with the CSR stored on disk gives the same GraphContext (bitwise) as dgdm_csr_build, and the
    on-device collate equals Batch.from_data_list."""
    from dgdm_histopath_lab_b200 import collate_to_device, load_graph, save_graph
    from dgdm_histopath_lab_b200.graph import get_context
    d = synth.tile_graph(2000, 48, seed=9)
    f = save_graph(d, str(tmp_path / "g.dgdm"))
    g = load_graph(f, device=DEV)
    ctx = g.__dict__["_dgdm_ctx"]
    ref = get_context(d.to(DEV))
    for k in ("rowptr", "col", "eid", "norm", "dinv"):
        assert torch.equal(getattr(ctx.csr, k)[: ref.csr.E if k in ("col", "eid", "norm") else None], getattr(ref.csr, k)[: ref.csr.E if k in ("col", "eid", "norm") else None]), k
    m = make_model(SMALL_CFG).to(DEV)
    a = m(g, return_embeddings=True)["node_embeddings"]
    assert g.__dict__["_dgdm_ctx"] is ctx, "the stored CSR must be used, not rebuilt"
    b = m(d.to(DEV), return_embeddings=True)["node_embeddings"]
    assert torch.equal(a, b)
    gs = [synth.tile_graph(k, 48, seed=40 + i) for i, k in enumerate((120, 77, 201))]
    for x in gs:
        x.edge_attr = torch.rand(x.edge_index.size(1), 2)
    cb = collate_to_device([x.pin_memory() for x in gs], DEV)
    rb = Batch.from_data_list(gs).to(DEV)
    for k in ("x", "edge_index", "pos", "edge_attr", "batch"):
        assert torch.equal(getattr(cb, k), getattr(rb, k)), k
    assert torch.equal(m(cb, return_embeddings=True)["node_embeddings"], m(rb, return_embeddings=True)["node_embeddings"])


@pytest.mark.parametrize("norm", ["batch", "instance"])
def test_non_default_normalization_full_model_vs_oracle(norm):
    cfg = dict(SMALL_CFG, normalization=norm)
    m = make_model(cfg)
    b = synth.tile_batch(2, 150, 48, seed=11)
    ref = O.dgdm_forward(sd_cpu(m), dict(oracle_cfg(cfg), normalization=norm), b.x, b.edge_index, None, b.batch, b.pos)
    out = m.to(DEV)(b.to(DEV), return_embeddings=True)
    assert rel_err(out["node_embeddings"], ref["node_embeddings"])[0] < TOL
    assert rel_err(out["graph_embedding"], ref["graph_embedding"])[0] < TOL


def test_diffusion_conditioning_vs_oracle():
    from dgdm_histopath_lab_b200 import DiffusionLayer
    torch.manual_seed(0)
    layer = DiffusionLayer(32, 64, num_timesteps=10, conditioning_dim=12)
    layer.load_state_dict(synth.seeded_state_dict(layer, 2))
    sd = {"d." + k: v.detach().clone() for k, v in layer.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x, c = torch.randn(50, 32, generator=g), torch.randn(50, 12, generator=g)
    t = torch.full((50,), 3, dtype=torch.long)
    ref = O.diffusion_predict_noise(sd, "d", x, t, c)
    out = layer.to(DEV).eval().predict_noise(x.to(DEV), t.to(DEV), c.to(DEV))
    assert rel_err(out, ref)[0] < 2e-5
    one = layer.predict_noise(x.to(DEV), t.to(DEV), c[:1].to(DEV))            # one condition row broadcast over the nodes
    assert rel_err(one, O.diffusion_predict_noise(sd, "d", x, t, c[:1].expand(50, -1)))[0] < 2e-5


def test_unsorted_batch_vector_is_accepted():
    """the reference selects nodes with boolean masks (models/dgdm_model.py:343-350), so `batch` may be in any order. (The model is
    NOT permutation-equivariant — D8 applies pooled-graph indices to the un-pooled node set — so the check is against the oracle
    on the same shuffled input, not against the sorted run.)"""
    m = make_model(SMALL_CFG)
    b = synth.tile_batch(3, 90, 48, seed=21)
    perm = torch.randperm(b.x.size(0), generator=torch.Generator().manual_seed(1))
    inv = torch.empty_like(perm); inv[perm] = torch.arange(perm.numel())
    sh = Data(x=b.x[perm], edge_index=inv[b.edge_index], pos=b.pos[perm])
    sh.batch = b.batch[perm]
    ref = O.dgdm_forward(sd_cpu(m), oracle_cfg(SMALL_CFG), sh.x, sh.edge_index, None, sh.batch, sh.pos)
    out = m.to(DEV)(sh.to(DEV), return_embeddings=True)
    assert rel_err(out["graph_embedding"], ref["graph_embedding"])[0] < TOL
    assert rel_err(out["node_embeddings"], ref["node_embeddings"])[0] < TOL
