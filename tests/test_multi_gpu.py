"""Real multi-GPU parity (one process per GPU, NCCL + symmetric memory). Launched by torchrun, e.g.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 -m pytest tests/test_multi_gpu.py -q -m gpu
Skipped in an ordinary single-process pytest run (the in-process ThreadComm test in test_gpu_model.py covers the
partition logic there)."""
import os

import pytest
import torch

from dgdm_histopath_lab_b200 import Data, synth
from util import make_model, rel_err

pytestmark = pytest.mark.gpu
WORLD = int(os.environ.get("WORLD_SIZE", "1"))


@pytest.mark.skipif(WORLD < 2, reason="needs torchrun with >= 2 ranks")
@pytest.mark.parametrize("fused", [False, True])
def test_partitioned_forward_over_nccl_matches_single_gpu(fused):
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(node_features=64, hidden_dims=[128, 64, 128], num_diffusion_steps=10, attention_heads=8, graph_layers=3, num_classes=2)
    n = 3001
    m = make_model(cfg).to(dev)
    d = synth.tile_graph(n, 64, seed=7).to(dev)
    ref = m(d, return_embeddings=True)
    m.set_distributed(rank, WORLD, max_nodes=n, fused=fused)
    if fused:
        assert type(m._comm).__name__ == "SymmComm", "symmetric memory / P2P should be available on an NVLink box"
    for _ in range(2):          # twice: the symmetric buffers are reused across forwards
        out = m(Data(x=d.x, edge_index=d.edge_index, pos=d.pos, num_nodes=n), return_embeddings=True)
    a, b = out["node_range"]
    assert rel_err(out["node_embeddings"], ref["node_embeddings"][a:b])[0] < 2e-5
    assert rel_err(out["graph_embedding"], ref["graph_embedding"])[0] < 2e-5
    m.set_partition(None)
    dist.barrier()
