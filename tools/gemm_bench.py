"""Times dgdm_gemm_f16x2 on the DGDM-Base GEMM shapes at M nodes (CUDA events, rotating buffers larger than L2) and prints the
HBM / tensor roofline fraction of each. Usage: python tools/gemm_bench.py [M] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
dev = "cuda"
g = torch.Generator().manual_seed(0)
# (K, N, epilogue) — the Base forward's shapes
shapes = [(768, 512, "ln_gelu"), (768, 512, "bias"), (512, 512, "ln_gelu_res"), (512, 512, "none"), (512, 512, "res_ln_ln_gelu"),
          (512, 256, "none"), (256, 256, "none"), (256, 512, "res_ln"), (512, 256, "ln_gelu"), (256, 128, "none"), (128, 128, "none"),
          (128, 256, "res_ln"), (128, 128, "res_ln_ln_gelu"), (128, 384, "bias"), (128, 64, "dot")]
if len(sys.argv) > 3:
    k_, n_, e_ = sys.argv[3].split(",")
    shapes = [(int(k_), int(n_), e_)]
nbuf = 4            # rotate inputs so that each launch reads cold data (4 x [M,768] fp32 = 1.2 GB > L2)
tot = 0.0
for K, N, ep in shapes:
    A = [torch.randn(M, K, generator=g).to(dev) for _ in range(nbuf)]
    W = ops.SplitWeight((torch.randn(N, K, generator=g) / K ** 0.5).to(dev))
    bias = torch.randn(N, generator=g).to(dev)
    ln = (torch.ones(N, device=dev), torch.zeros(N, device=dev))
    res = torch.randn(M, N, generator=g).to(dev)
    kw = {}
    extra = 0
    if ep == "bias":
        kw = dict(bias=bias)
    elif ep == "ln_gelu":
        kw = dict(bias=bias, ln1=ln, act_post="gelu")
    elif ep == "ln_gelu_res":
        kw = dict(bias=bias, ln1=ln, act_post="gelu", r_post=res); extra = M * N * 4
    elif ep == "res_ln":
        kw = dict(bias=bias, r_pre=res, ln1=ln); extra = M * N * 4
    elif ep == "res_ln_ln_gelu":
        kw = dict(bias=bias, r_pre=res, ln1=ln, ln2=ln, act_post="gelu"); extra = M * N * 4
    elif ep == "dot":
        kw = dict(bias=bias, act_post="relu", want_out=False, dot=(bias, 0.1, "tanh"))
    out = torch.empty(M, N, device=dev)
    for i in range(3):
        ops.linear(A[i % nbuf], W, out=out if ep != "dot" else None, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.linear(A[i % nbuf], W, out=out if ep != "dot" else None, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = M * K * 4 + (M * N * 4 if ep != "dot" else M * 4) + extra
    fl = 3 * 2.0 * M * K * N
    tot += ms
    print(f"K={K:4d} N={N:4d} {ep:16s} {ms*1e3:8.1f} us   {byts/ms/1e6:7.0f} GB/s ({byts/ms/1e6/pk['hbm_gbs']*100:5.1f}% hbm)   "
          f"{fl/ms/1e9:7.1f} TF/s x3 ({fl/ms/1e9/pk['bf16_tflops']*100:5.1f}% tensor)", flush=True)
    del A, res, out
print(f"sum {tot:.3f} ms")
