import os
import warnings

import numpy as np
import torch

from dgdm_histopath_lab_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_CFG = dict(node_features=48, hidden_dims=[64, 32, 32], num_diffusion_steps=10, attention_heads=2, graph_layers=4,
                 num_classes=3, regression_targets=2)
BUILDER_MODEL_CFG = dict(node_features=64, hidden_dims=[64, 32, 32], num_diffusion_steps=10, attention_heads=2, graph_layers=3,
                         num_classes=3)
BASE_CFG = dict(node_features=768, hidden_dims=[512, 256, 128], num_diffusion_steps=10, attention_heads=8, graph_layers=4)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def make_model(cfg, seed=0, device="cpu"):
    from dgdm_histopath_lab_b200 import DGDMModel
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = DGDMModel(**cfg).eval()
    m.load_state_dict(synth.seeded_state_dict(m, seed))
    return m.to(device)


def sd_cpu(model):
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}


def oracle_cfg(cfg):
    return dict(hidden_dims=cfg["hidden_dims"], attention_heads=cfg["attention_heads"], graph_layers=cfg["graph_layers"],
                activation="gelu", use_spatial_attention=True, use_hierarchical=True,
                num_diffusion_steps=cfg["num_diffusion_steps"], diffusion_schedule="cosine")


def rel_err(a, b):
    """max |a-b| / max|b| (scale-relative) and the row-wise relative L2 error (max over rows)."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64) if not torch.is_tensor(a) else a.detach().cpu().double()
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64) if not torch.is_tensor(b) else b.detach().cpu().double()
    scale = b.abs().max().clamp_min(1e-30)
    mx = ((a - b).abs().max() / scale).item()
    if a.dim() >= 2:
        num = (a - b).reshape(a.shape[0], -1).norm(dim=1)
        den = b.reshape(b.shape[0], -1).norm(dim=1).clamp_min(1e-30)
        row = (num / den).max().item()
    else:
        row = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    return mx, row


def pretrain_draws(seed, sizes, dim, steps):
    """Order of torch RNG calls in the reference's _compute_diffusion_loss (models/dgdm_model.py:411-429)."""
    torch.manual_seed(seed)
    t = torch.randint(0, steps, (len(sizes),))
    noise, target = [], []
    for n in sizes:
        noise.append(torch.randn(1, n, dim)[0])
        target.append(torch.randn(n, dim))
    return t, noise, target
