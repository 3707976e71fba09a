"""One spatial-attention call at n nodes (for ncu captures). Usage: python tools/attn_only.py [n] [precise] [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = torch.Generator().manual_seed(0)
qkv = torch.randn(n, 384, generator=g).cuda(); pos = torch.rand(n, 2, generator=g).cuda()
seg = torch.tensor([0, n], dtype=torch.int32, device="cuda")
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); o = ops.spatial_attention(qkv, pos, seg, 1, 8, 1.0, precise=prec); b.record(); torch.cuda.synchronize()
    print(f"attention n={n} precise={prec}: {a.elapsed_time(b):.2f} ms  checksum {o.double().abs().sum().item():.6f}")
