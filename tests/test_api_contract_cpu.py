"""CPU: the host-side contract of the drop-in classes — constructor validation and error types as pinned by the reference's own
API tests (dgdm_histopath/testing/test_suite.py:51-115), the graph containers, the diffusion schedule tables, and the guarantee that
the product path has NO CPU fallback (CPU tensors raise instead of silently computing in torch)."""
import warnings

import pytest
import torch

from dgdm_histopath_lab_b200 import (Batch, Data, DGDMModel, DiffusionScheduler, GraphConvolution, ModelConfigurationError,
                                     ModelInferenceError, synth)
from util import SMALL_CFG, make_model


@pytest.mark.parametrize("bad", [
    dict(node_features=0), dict(node_features=20000), dict(hidden_dims=[]), dict(hidden_dims=[64, -1]), dict(attention_heads=0),
    dict(attention_heads=33), dict(hidden_dims=[64, 30], attention_heads=4), dict(dropout=0.95), dict(graph_layers=0), dict(graph_layers=21),
    dict(activation="nope"), dict(pooling="nope"), dict(diffusion_schedule="nope"), dict(normalization="nope"), dict(num_classes=1),
])
def test_constructor_rejects_invalid_configuration(bad):
    cfg = dict(node_features=48, hidden_dims=[64, 32, 32], attention_heads=2, num_classes=3)
    cfg.update(bad)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(ModelConfigurationError):
            DGDMModel(**cfg)


def test_constructor_warns_without_heads_and_keeps_reference_attributes():
    with pytest.warns(UserWarning, match="No classification or regression targets"):
        m = DGDMModel(node_features=48, hidden_dims=[64, 32, 32], attention_heads=2)
    assert m.node_features == 48 and m.hidden_dims == [64, 32, 32] and m.num_classes is None
    assert m.classification_head is None and m.regression_head is None and m.num_diffusion_steps == 10
    # dead parameters of the reference stay in the state_dict (D14 / D15) so that reference checkpoints load unchanged
    keys = m.state_dict().keys()
    assert "spatial_attention.pos_encoding" in keys and "graph_encoder.graph_layers.0.norm2.weight" in keys
    assert "global_pool.global_token" in keys and "graph_encoder.graph_layers.0.node_to_qkv.weight" in keys
    rt = DGDMModel(node_features=48, hidden_dims=[64, 32, 32], attention_heads=2, num_classes=None)
    rt.load_state_dict(m.state_dict())                               # round trip (testing/robust_validators.py:682-705)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), rt.state_dict().values()))


def test_no_cpu_fallback_forward_raises_model_inference_error():
    m = make_model(SMALL_CFG)
    d = synth.tile_graph(50, 48, seed=1)
    with pytest.raises(ModelInferenceError):
        m(d)                                                         # CPU tensors: refuse, do not compute
    with pytest.raises(Exception):
        GraphConvolution(8, 8)(torch.randn(5, 8), torch.zeros(2, 0, dtype=torch.long))


def test_forward_rejects_malformed_containers_before_touching_the_device():
    m = make_model(SMALL_CFG)
    with pytest.raises(ModelInferenceError):
        m({"x": torch.zeros(3, 48)})                                 # not a Data / Batch
    with pytest.raises(ModelInferenceError):
        m(Data(x=torch.zeros(3, 48)))                                # missing edge_index
    with pytest.raises(ModelInferenceError):
        m(Data(x=torch.zeros(3, 7), edge_index=torch.zeros(2, 0, dtype=torch.long)))      # wrong feature width
    with pytest.raises(ModelInferenceError):
        m(synth.tile_graph(20, 48, seed=1), mode="nope")


def test_batch_container_matches_pyg_semantics():
    a, b = synth.tile_graph(5, 4, seed=1), synth.tile_graph(7, 4, seed=2)
    bt = Batch.from_data_list([a, b])
    assert bt.x.shape == (12, 4) and bt.num_graphs == 2
    assert torch.equal(bt.batch, torch.tensor([0] * 5 + [1] * 7))
    assert torch.equal(bt.edge_index[:, a.edge_index.size(1):], b.edge_index + 5)
    assert torch.equal(bt.pos[5:], b.pos)
    c = bt.clone()
    c.x[0, 0] += 1
    assert not torch.equal(c.x, bt.x) and c.num_nodes == 12


def test_diffusion_schedule_tables_match_closed_forms():
    s = DiffusionScheduler(10, schedule="cosine")
    assert s.betas.shape == (10,) and float(s.betas.max()) <= 0.999 + 1e-6 and float(s.betas.min()) >= 0
    assert torch.allclose(s.alphas_cumprod, torch.cumprod(1 - s.betas, 0))
    assert float(s.alphas_cumprod_prev[0]) == 1.0
    lin = DiffusionScheduler(5, beta_start=0.1, beta_end=0.5, schedule="linear")
    assert torch.allclose(lin.betas, torch.linspace(0.1, 0.5, 5))
    with pytest.raises(ValueError):
        DiffusionScheduler(5, schedule="nope")


@pytest.mark.parametrize("norm,cls", [("batch", torch.nn.BatchNorm1d), ("instance", torch.nn.InstanceNorm1d), ("graph", torch.nn.Identity)])
def test_non_default_normalizations_construct_like_the_reference(norm, cls):
    """models/encoders.py:93-102, 217-226: batch / instance build the torch layers, any other accepted name ("graph") is Identity"""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = DGDMModel(node_features=48, hidden_dims=[64, 32, 32], attention_heads=2, num_classes=3, normalization=norm)
    assert isinstance(m.feature_encoder.encoder[1], cls) and isinstance(m.graph_encoder.norm_layers[0], cls)


def test_validation_result_interpretation_matches_the_reference_messages():
    """Host half of the one-kernel input validation (models/dgdm_model.py:646-710): the 4 flags map to the reference's error texts in
    its order of checks, the statistics to its warnings. The same function serves the immediate read-back of `forward` and the
    deferred one of `graphed(validate=True)` / `loader.GraphServer`."""
    from dgdm_histopath_lab_b200.errors import ValidationError
    chk = DGDMModel._check_validation_result
    ok_stats = [1.0, 10.0, 200.0, 0.0]                       # max |x|, sum, sum of squares
    chk([0, 0, 0, 0], ok_stats, 100)
    for flags, msg in (([1, 1, 1, 1], "NaN"), ([0, 1, 1, 1], "infinity"), ([0, 0, 1, 1], "invalid node indices"),
                       ([0, 0, 0, 1], "negative node indices")):
        with pytest.raises(ValidationError, match=msg):
            chk(flags, ok_stats, 100)
    with pytest.warns(UserWarning, match="Large feature values"):
        chk([0, 0, 0, 0], [5000.0, 10.0, 200.0, 0.0], 100)
    with pytest.warns(UserWarning, match="very low variance"):
        chk([0, 0, 0, 0], [1.0, 100.0, 100.0, 0.0], 100)     # all ones: variance 0
