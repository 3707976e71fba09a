"""Multi-GPU microbench (torchrun): cost of the pieces of the fused halo exchange — the signal-pad barrier, the GEMM with
peer / multicast stores, the plain GEMM, and NCCL all_gather of the same matrix."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import ops
from dgdm_histopath_lab_b200.parallel import SymmComm, Part, even_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
comm = SymmComm(N * 512, dev)
part = Part(comm, even_bounds(N, world))

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

res = {}
res["barrier_us"] = timed(lambda: comm.hdls[0].barrier(channel=0), 50)
for K, C in [(512, 512), (128, 128)]:
    x = torch.randn(part.n_local, K, device=dev); W = ops.SplitWeight(torch.randn(C, K, device=dev) / K ** 0.5)
    res[f"gemm_local_{K}x{C}_us"] = timed(lambda: ops.linear(x, W))
    for mc in ([True, False] if comm.multicast else [False]):
        comm.multicast = mc
        res[f"gemm_fused_{'mc' if mc else 'p2p'}_{K}x{C}_us"] = timed(lambda: comm.fused_linear_gather(lambda out, peers, m: ops.linear(x, W, out=out, peers=peers, multicast=m), part, C))
    h = ops.linear(x, W)
    res[f"nccl_allgather_{C}_us"] = timed(lambda: part.gather(h))
    full_mb = N * C * 4 / 1e6
    res[f"matrix_MB_{C}"] = full_mb
if rank == 0:
    print(world, "GPUs, N =", N)
    for k, v in res.items(): print(f"  {k:32s} {v:10.1f}")
dist.destroy_process_group()
