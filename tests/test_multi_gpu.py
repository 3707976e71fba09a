"""Real multi-PROCESS parity of the node-partitioned forward (one process per rank, torch.distributed): the test spawns
`python -m torch.distributed.run --nproc-per-node 2 tests/mgpu_worker.py` itself, so the driver's plain
`pytest -m gpu` run exercises it. With >= 2 GPUs: NCCL + symmetric memory (P2P halo stores fused into the GEMM epilogue);
on a single-GPU box: two processes share the GPU over a gloo group (see mgpu_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_partitioned_forward_two_processes_matches_single_gpu(world):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None); env.pop("LOCAL_RANK", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "mgpu_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "MGPU_WORKER_OK" in r.stdout
    if torch.cuda.device_count() >= world:
        assert "gpu_per_rank" in r.stdout
