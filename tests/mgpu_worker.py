"""Worker for tests/test_multi_gpu.py: one process per rank (torch.distributed.run). With >= WORLD_SIZE GPUs every rank owns
a GPU and the group is NCCL; on a single-GPU box all ranks share cuda:0 and the group is gloo (NCCL refuses two ranks on one
device) — the partition logic, TorchDistComm collectives and, where CUDA IPC allows it, the symmetric-memory fused halo stores
(SymmComm) still run as separate processes. Exit code 0 = parity with the single-GPU forward."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from dgdm_histopath_lab_b200 import Data, synth          # noqa: E402
from util import make_model, rel_err                       # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    shared = ngpu < world
    dev = torch.device("cuda", 0 if shared else int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(node_features=64, hidden_dims=[128, 64, 128], num_diffusion_steps=10, attention_heads=8, graph_layers=3, num_classes=2)
    n = 3001
    m = make_model(cfg).to(dev)
    d = synth.tile_graph(n, 64, seed=7).to(dev)
    ea = torch.randn(d.edge_index.size(1), 2, generator=torch.Generator().manual_seed(1)).to(dev)
    report = []
    for with_attr in (False, True):
        m.set_partition(None)
        attr = ea if with_attr else None
        ref = m(Data(x=d.x, edge_index=d.edge_index, edge_attr=attr, pos=d.pos), return_embeddings=True)
        for fused in (False, True):
            m.set_distributed(rank, world, max_nodes=n, fused=fused)
            kind = type(m._comm).__name__
            if fused and not shared:
                assert kind == "SymmComm", "symmetric memory / P2P should be available on an NVLink box"
            for _ in range(2):          # twice: the symmetric buffers are reused across forwards
                out = m(Data(x=d.x, edge_index=d.edge_index, edge_attr=attr, pos=d.pos, num_nodes=n), return_embeddings=True)
            a, b = out["node_range"]
            e1 = rel_err(out["node_embeddings"], ref["node_embeddings"][a:b])[0]
            e2 = rel_err(out["graph_embedding"], ref["graph_embedding"])[0]
            assert e1 < 2e-5 and e2 < 2e-5, (with_attr, fused, kind, e1, e2)
            report.append((with_attr, kind, e1))
            dist.barrier()
    m.set_partition(None)
    if rank == 0:
        print("MGPU_WORKER_OK", "shared_gpu" if shared else "gpu_per_rank", report, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
