"""CPU: the oracle restatement (oracle/dgdm_oracle.py, oracle/builder_oracle.py) against the golden vectors that
oracle/make_golden.py produced by running the real (patched) reference. Tolerance: fp32 re-association only."""
import numpy as np
import torch

from dgdm_histopath_lab_b200 import synth
from oracle import builder_oracle, dgdm_oracle as O
from util import BASE_CFG, SMALL_CFG, golden, make_model, oracle_cfg, pretrain_draws, rel_err, sd_cpu

TOL = 2e-5


def test_small_model_inference_and_state_dict_keys():
    g = golden("small_model.npz")
    m = make_model(SMALL_CFG)
    assert sorted(m.state_dict().keys()) == list(g["state_keys"]), "state_dict key layout differs from the reference"
    b = synth.tile_batch(2, 150, SMALL_CFG["node_features"], seed=11)
    out = O.dgdm_forward(sd_cpu(m), oracle_cfg(SMALL_CFG), b.x, b.edge_index, None, b.batch, b.pos, return_stages=True)
    for k, ref in (("feature_encoder", g["feature_encoder"]), ("graph_encoder", g["graph_encoder"])):
        assert rel_err(out["stages"][k], ref)[0] < TOL, k
    assert rel_err(out["node_embeddings"], g["node_embeddings"])[0] < 1e-4
    assert rel_err(out["graph_embedding"], g["graph_embedding"])[0] < 1e-4
    assert rel_err(out["classification_logits"], g["classification_logits"])[0] < 1e-4


def test_small_model_pretrain_replayed_rng():
    g = golden("small_model.npz")
    m = make_model(SMALL_CFG)
    b = synth.tile_batch(2, 150, SMALL_CFG["node_features"], seed=11)
    t, noise, target = pretrain_draws(int(g["pretrain_seed"]), [150, 150], 32, 10)
    out = O.dgdm_forward(sd_cpu(m), oracle_cfg(SMALL_CFG), b.x, b.edge_index, None, b.batch, b.pos, mode="pretrain",
                         rng=dict(timesteps=t, noise=noise, target=target))
    assert abs(float(out["diffusion_loss"]) - float(g["diffusion_loss"])) < 1e-4 * abs(float(g["diffusion_loss"]))
    assert rel_err(out["noisy_embeddings"], g["noisy_embeddings"])[0] < 1e-4


def test_edge_attr_encoder_and_dynamic_edges():
    g = golden("small_edge_attr.npz")
    m = make_model(SMALL_CFG)
    sd = sd_cpu(m)
    b = synth.tile_batch(2, 150, SMALL_CFG["node_features"], seed=11)
    ea = torch.randn(b.edge_index.size(1), 32, generator=torch.Generator().manual_seed(5))
    fe = O.feature_encoder(sd, "feature_encoder", b.x)
    enc = O.graph_encoder(sd, "graph_encoder", fe, b.edge_index, ea, 4)
    assert rel_err(enc, g["graph_encoder"])[0] < TOL
    dyn = O.compute_dynamic_edges(sd, "graph_encoder.graph_layers.1", fe, b.edge_index, ea, 2)
    assert rel_err(dyn, g["dynamic_edges"])[0] < TOL


def test_diffusion_and_pooling_modules():
    g = golden("modules.npz")
    m = make_model(SMALL_CFG)
    sd = sd_cpu(m)
    sched = O.diffusion_schedule(10, "cosine")
    t = torch.full((200,), 4, dtype=torch.long)
    xn, eps = O.diffusion_forward(sd, "diffusion_layer", sched, torch.tensor(g["x0"]), t, torch.tensor(g["noise"]))
    assert rel_err(xn, g["x_noisy"])[0] < 1e-6 and rel_err(eps, g["eps"])[0] < TOL
    s = O.diffusion_sample(sd, "diffusion_layer", sched, torch.tensor(g["sample_init"]), torch.tensor(g["sample_steps"]), 6)
    assert rel_err(s, g["sample_out"])[0] < 1e-4
    gm = golden("small_model.npz")
    b = synth.tile_batch(2, 150, SMALL_CFG["node_features"], seed=11)
    px, pei, _, perm, _ = O.adaptive_graph_pooling(sd, "hierarchical_processor.pools.0", torch.tensor(gm["graph_encoder"]), b.edge_index)
    assert np.array_equal(perm.numpy(), g["pool_perm"]) and np.array_equal(pei.numpy(), g["pool_ei"])
    assert rel_err(px, g["pool_x"])[0] < TOL


def test_base_config_2k_nodes():
    g = golden("base_2k.npz")
    m = make_model(BASE_CFG)
    assert sum(p.numel() for p in m.parameters()) == int(g["num_params"])
    d = synth.tile_graph(2000, 768, seed=1234)
    out = O.dgdm_forward(sd_cpu(m), oracle_cfg(BASE_CFG), d.x, d.edge_index, None, None, d.pos, return_stages=True)
    assert rel_err(out["stages"]["graph_encoder"][::8], g["graph_encoder"])[0] < TOL
    assert rel_err(out["stages"]["spatial_attention"][::8], g["spatial"])[0] < TOL
    assert rel_err(out["node_embeddings"], g["node_embeddings"])[0] < 1e-4
    assert rel_err(out["graph_embedding"], g["graph_embedding"])[0] < 1e-4


def test_builder_edges_exact():
    g = golden("builder_600.npz")
    n = 600
    feats = synth.clustered_features(n, d=64, num_clusters=12, seed=21, sigma_lo=0.05, sigma_hi=0.11)
    xs, ys, dims = synth.jittered_grid_coords(n, seed=21, extent=3000)
    coords = np.stack([xs / dims[0], ys / dims[1]], axis=1)
    ei, ea, et = builder_oracle.build_edges(feats, coords)
    assert np.array_equal(ei, g["edge_index"]) and np.array_equal(et, g["edge_type"])
    assert np.allclose(ea, g["edge_attr"], rtol=0, atol=1e-7)


# ------------------------------------------------------------------------------------------ round-2 fixtures
def _edge_attrs(b):
    g = torch.Generator().manual_seed(5)
    return torch.randn(b.edge_index.size(1), 32, generator=g), torch.rand(b.edge_index.size(1), 2, generator=g), g


def test_full_model_with_edge_attr_through_unet():
    """edge attributes through every GraphUNet level (core/graph_layers.py:404-453, pooled by edge_attr[edge_mask] :327)"""
    g = golden("edge_attr_model.npz")
    m = make_model(SMALL_CFG)
    sd = sd_cpu(m)
    b = synth.tile_batch(2, 150, SMALL_CFG["node_features"], seed=11)
    ea32, ea2, gen = _edge_attrs(b)
    for tag, ea in (("ea32", ea32), ("ea2", ea2)):
        out = O.dgdm_forward(sd, oracle_cfg(SMALL_CFG), b.x, b.edge_index, ea, b.batch, b.pos)
        assert rel_err(out["node_embeddings"], g[tag + "_node_embeddings"])[0] < 1e-4, tag
        assert rel_err(out["graph_embedding"], g[tag + "_graph_embedding"])[0] < 1e-4, tag
        assert rel_err(out["classification_logits"], g[tag + "_logits"])[0] < 1e-4, tag
    d1 = synth.tile_graph(257, 32, seed=5)
    ea1 = torch.randn(d1.edge_index.size(1), 32, generator=gen)
    u, perms = O.graph_unet(sd, "hierarchical_processor", d1.x, d1.edge_index, ea1, return_perms=True)
    for i, pm in enumerate(perms):
        assert np.array_equal(pm.numpy(), g[f"unet1_perm{i}"]), f"perm level {i}"
    assert rel_err(u, g["unet1_out"])[0] < TOL


def builder_graph_600():
    n = 600
    feats = synth.clustered_features(n, d=64, num_clusters=12, seed=21, sigma_lo=0.05, sigma_hi=0.11)
    xs, ys, dims = synth.jittered_grid_coords(n, seed=21, extent=3000)
    return feats, xs, ys, dims


def test_build_graph_then_forward():
    """the reference's serving pipeline: TissueGraphBuilder.build_graph -> DGDMModel.forward (evaluation/predictor.py:188-215)"""
    from util import BUILDER_MODEL_CFG
    gb, gf = golden("builder_600.npz"), golden("builder_forward_600.npz")
    feats, xs, ys, dims = builder_graph_600()
    m = make_model(BUILDER_MODEL_CFG, seed=3)
    out = O.dgdm_forward(sd_cpu(m), oracle_cfg(BUILDER_MODEL_CFG), torch.tensor(feats), torch.tensor(gb["edge_index"]),
                         torch.tensor(gb["edge_attr"]), None, torch.tensor(gb["pos"]))
    assert rel_err(out["node_embeddings"], gf["node_embeddings"])[0] < 1e-4
    assert rel_err(out["graph_embedding"], gf["graph_embedding"])[0] < 1e-4
    assert rel_err(out["classification_logits"], gf["logits"])[0] < 1e-4


def test_pretrain_step_replayed_rng():
    g = golden("pretrain_step.npz")
    m = make_model(SMALL_CFG)
    b = synth.tile_batch(2, 150, SMALL_CFG["node_features"], seed=11)
    rng = dict(mask_indices=torch.tensor(g["mask_indices"]), mask_token=torch.tensor(g["mask_token"]),
               timesteps=torch.tensor(g["timesteps"]), noise=list(torch.tensor(g["noise"]).split(150)),
               target=list(torch.tensor(g["target"]).split(150)))
    out = O.pretrain_step(sd_cpu(m), oracle_cfg(SMALL_CFG), b.x, b.edge_index, None, b.batch, b.pos, rng=rng)
    assert abs(float(out["total_pretrain_loss"]) - float(g["total_pretrain_loss"])) < 1e-4 * abs(float(g["total_pretrain_loss"]))
    assert rel_err(out["noisy_embeddings"], g["noisy_embeddings"])[0] < 1e-4


def test_public_attention_api():
    g = golden("mha_public.npz")
    m = make_model(SMALL_CFG)
    sd = sd_cpu(m)
    p = "spatial_attention.attention"
    q, k, v = (torch.tensor(g[n]) for n in ("q", "k", "v"))
    o, w = O.multi_head_attention(sd, p, q, num_heads=2)
    assert rel_err(o, g["self_out"])[0] < TOL and rel_err(w, g["self_w"])[0] < TOL
    o, w = O.multi_head_attention(sd, p, q, num_heads=2, attn_mask=torch.tensor(g["attn_mask_sq"]), average_attn_weights=False)
    assert w.shape == (4, 37, 37)
    assert rel_err(o, g["selfmask_out"])[0] < TOL and rel_err(w, g["selfmask_w"])[0] < TOL
    o, w = O.multi_head_attention(sd, p, q, k, v, num_heads=2, attn_mask=torch.tensor(g["attn_mask"]),
                                  key_padding_mask=torch.tensor(g["key_padding_mask"]))
    assert rel_err(o, g["cross_out"])[0] < TOL and rel_err(w, g["cross_w"])[0] < TOL
    o, _ = O.multi_head_attention(sd, p, q, k, num_heads=2, need_weights=False)
    assert rel_err(o, g["cross_kv_out"])[0] < TOL
    o, w = O.spatial_attention_batched(sd, "spatial_attention", torch.tensor(g["spatial_x"]), torch.tensor(g["spatial_pos"]), 2)
    assert rel_err(o, g["spatial_out"])[0] < TOL and rel_err(w, g["spatial_w"])[0] < TOL
