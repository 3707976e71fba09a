"""Kernel-level profile of one pretrain step (torch.profiler, CUDA activities): python tools/train_profile.py [nodes]"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import DGDMModel, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = DGDMModel(node_features=768, hidden_dims=[512, 256, 128], num_diffusion_steps=10, attention_heads=8, graph_layers=4)
m.load_state_dict(synth.seeded_state_dict(m, 0))
m = m.cuda().train()
d = synth.tile_graph(n, 768, seed=500).cuda()


def step():
    m.zero_grad(set_to_none=True)
    out = m.pretrain_step(d, mask_ratio=0.15)
    out["total_pretrain_loss"].backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
