"""Micro-timings on the GPU box (CUDA events, warm, median of reps). Usage: python tools/probe.py [gemm] [attn] [spmm]"""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import ops, synth

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

what = sys.argv[1:] or ["gemm", "attn"]
dev = "cuda"
res = {}
if "gemm" in what:
    for (M, K, N) in [(1, 128, 128), (128, 128, 128), (18944, 128, 128), (100000, 128, 128), (100000, 768, 512), (100000, 512, 512), (100000, 128, 384)]:
        a = torch.randn(M, K, device=dev); w = ops.SplitWeight(torch.randn(N, K, device=dev) / K ** 0.5)
        out = torch.empty(M, N, device=dev)
        t = timeit(lambda: ops.linear(a, w, out=out))
        gb = (M * K + M * N) * 4 / 1e9
        res[f"gemm_{M}x{K}x{N}"] = dict(ms=t, GBps=gb / (t * 1e-3), TFLOPs_3pass=3 * 2 * M * K * N / (t * 1e-3) / 1e12)
        print(f"gemm M={M} K={K} N={N}: {t*1e3:.1f} us  {gb/(t*1e-3):.0f} GB/s algorithmic", flush=True)
    M, K, N = 100000, 512, 512
    a = torch.randn(M, K, device=dev); w = ops.SplitWeight(torch.randn(N, K, device=dev) / K ** 0.5)
    g = torch.ones(N, device=dev); b = torch.zeros(N, device=dev); r = torch.randn(M, N, device=dev)
    t = timeit(lambda: ops.linear(a, w, b, r_pre=r, ln1=(g, b), ln2=(g, b), act_post="gelu"))
    print(f"gemm+LN+LN+gelu M={M} K={K} N={N}: {t*1e3:.1f} us", flush=True)
    res["gemm_ln2_100000x512x512"] = dict(ms=t)
if "attn" in what:
    for n in (20000, 100000):
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(n, 384, generator=g).to(dev); pos = torch.rand(n, 2, generator=g).to(dev)
        seg = torch.tensor([0, n], dtype=torch.int32, device=dev)
        for prec in (1, 0):
            t = timeit(lambda: ops.spatial_attention(qkv, pos, seg, 1, 8, 1.0, precise=prec), reps=5, warm=2)
            res[f"attn_n{n}_precise{prec}"] = dict(ms=t, gexp_per_s=8.0 * n * n / (t * 1e-3) / 1e9)
            print(f"attention n={n} precise={prec}: {t:.2f} ms  {8.0*n*n/(t*1e-3)/1e9:.0f} Gexp/s", flush=True)
        o1 = ops.spatial_attention(qkv, pos, seg, 1, 8, 1.0, precise=1); o0 = ops.spatial_attention(qkv, pos, seg, 1, 8, 1.0, precise=0)
        print(f"  fast-vs-precise rel diff (max/absmax): {((o1-o0).abs().max()/o1.abs().max()).item():.3e}", flush=True)
if "spmm" in what:
    d = synth.tile_graph(100000, 8, seed=1234)
    csr = ops.csr_build(d.edge_index.to(dev), 100000)
    for C in (128, 256, 512):
        h = torch.randn(100000, C, device=dev); out = torch.empty_like(h)
        t = timeit(lambda: ops.spmm_gcn(csr, h, out=out, act="gelu"))
        E = d.edge_index.size(1); by = (E + 100000) * C * 4 + 100000 * C * 4 + (E + 100000) * 8
        print(f"spmm C={C}: {t*1e3:.1f} us  {by/(t*1e-3)/1e9:.0f} GB/s algorithmic", flush=True)
        res[f"spmm_C{C}"] = dict(ms=t, GBps=by / (t * 1e-3) / 1e9)
    t = timeit(lambda: ops.csr_build(d.edge_index.to(dev), 100000), reps=5)
    print(f"csr_build 100k/1.6M: {t*1e3:.1f} us"); res["csr_build"] = dict(ms=t)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
