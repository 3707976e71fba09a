"""Dev tool: CTA-0 event timeline of gemm_f16x2_kernel (library built with -DDGDM_GEMM_TRACE by tools/gemm_trace.sh)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import ops, _lib
lib = _lib.load()
M = 100000
names = {11: "prologue done", 4: "conv: slot free", 5: "conv: body done", 6: "conv: fence done", 0: "conv: arrived", 7: "prod: w slot free",
         8: "mma: full_a seen", 9: "mma: full_w seen", 10: "mma: item issued", 1: "mma: tile commit", 2: "epi: start", 3: "epi: end"}
shapes = [(128, 128), (512, 512)] if len(sys.argv) < 2 else [tuple(int(v) for v in sys.argv[1].split(","))]
for K, N in shapes:
    a = torch.randn(M, K, device="cuda"); W = ops.SplitWeight(torch.randn(N, K, device="cuda") / K ** 0.5)
    ops.linear(a, W); torch.cuda.synchronize()
    buf = (C.c_ulonglong * (12 * 64))()
    lib.dgdm_debug_gemm_trace(None, 1)
    ops.linear(a, W); torch.cuda.synchronize()
    lib.dgdm_debug_gemm_trace(buf, 0)
    ev = [list(buf)[64 * r: 64 * (r + 1)] for r in range(12)]
    print(f"K={K} N={N}: CTA0 events, cycles since kernel start")
    for r in (11, 4, 5, 6, 0, 7, 8, 9, 10, 1, 2, 3):
        print(f"   {names[r]:20s}", [v for v in ev[r][:14]])
