"""Attention forward / backward timing (training path) at n nodes: python tools/attn_bwd_time.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
g = torch.Generator().manual_seed(0)
qkv = torch.randn(n, 384, generator=g).cuda()
pos = torch.rand(n, 2, generator=g).cuda()
go = torch.randn(n, 128, generator=g).cuda()
seg = torch.tensor([0, n], dtype=torch.int32, device="cuda")
seg_of = torch.zeros(n, dtype=torch.int32, device="cuda")
o, lse = ops.spatial_attention(qkv, pos, seg, 1, 8, want_lse=True)
ops.spatial_attention_backward(qkv, pos, seg_of, seg, 8, 1.0, o, go, lse=lse)
o, lse = ops.spatial_attention(qkv, pos, seg, 1, 8, want_lse=True)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
o, lse = ops.spatial_attention(qkv, pos, seg, 1, 8, want_lse=True)
ev[1].record()
d = ops.spatial_attention_backward(qkv, pos, seg_of, seg, 8, 1.0, o, go, lse=lse)
ev[2].record()
torch.cuda.synchronize()
fw, bw = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
pairs = float(n) * n * 8
print(f"n={n}: forward {fw:.2f} ms, backward {bw:.2f} ms ({bw / fw:.1f}x); backward {pairs * 128 * 2 / bw / 1e9:.1f} TFLOP/s fp32 (128 FMA per score)")
