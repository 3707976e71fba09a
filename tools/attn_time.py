"""Attention-only timing at the BASELINE size (100k nodes, 8 heads x 16) + a quick accuracy probe against float64 on sampled rows.
    python tools/attn_time.py [nodes] [reps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda"
g = torch.Generator().manual_seed(3)
qkv = torch.randn(n, 384, generator=g).to(dev)
pos = torch.rand(n, 2, generator=g).to(dev)
seg = torch.tensor([0, n], dtype=torch.int32, device=dev)
for _ in range(2):
    o = ops.spatial_attention(qkv, pos, seg, 1, 8)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    o = ops.spatial_attention(qkv, pos, seg, 1, 8)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
# accuracy: 64 sampled rows vs float64
idx = torch.randint(0, n, (64,), generator=g).to(dev)
q = qkv[idx, :128].double().view(64, 8, 16)
k = qkv[:, 128:256].double().view(n, 8, 16)
v = qkv[:, 256:].double().view(n, 8, 16)
s = torch.einsum("qhd,khd->hqk", q, k) / 4.0
s = s - torch.cdist(pos[idx].double(), pos.double()).unsqueeze(0)
w = torch.softmax(s, dim=-1)
ref = torch.einsum("hqk,khd->qhd", w, v).reshape(64, 128)
err = ((o[idx].double() - ref).abs().max() / ref.abs().max()).item()
print(f"attention {n} nodes: {ms:.3f} ms  variant={os.environ.get('DGDM_ATTN_TC_WARPS', 'default')}  rel err vs fp64 (64 rows) {err:.2e}")
