"""Host-side logic of the node-partitioned execution (dgdm_histopath_lab_b200/parallel.py) on CPU: partition bookkeeping and
the row all-gather over a real 2-rank torch.distributed group (gloo). No CUDA kernel is called here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dgdm_histopath_lab_b200.parallel import Part, ThreadComm, TorchDistComm, even_bounds


def test_even_bounds_cover_all_rows_and_handle_short_tails():
    assert even_bounds(100000, 8) == [12500 * r for r in range(9)]
    assert even_bounds(10, 4) == [0, 3, 6, 9, 10]
    assert even_bounds(2, 4) == [0, 1, 2, 2, 2]          # trailing ranks own nothing
    assert even_bounds(0, 2) == [0, 0, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = TorchDistComm()
        ok = True
        # even blocks, 2-D rows
        full = torch.arange(24, dtype=torch.float32).view(8, 3)
        p = Part(comm, even_bounds(8, world))
        ok &= torch.equal(p.gather(p.local(full).clone()), full)
        # uneven blocks (a pooled U-Net level) incl. 1-D tensors (scores)
        p2 = Part(comm, [0, 5, 8])
        ok &= torch.equal(p2.gather(p2.local(full).clone()), full)
        s = torch.arange(8, dtype=torch.float32) * 0.5
        ok &= torch.equal(p2.gather(p2.local(s).clone()), s)
        # pooled partition: node_map is replicated, a kept node stays on its owner
        node_map = torch.tensor([0, -1, 1, 2, -1, -1, 3, -1], dtype=torch.int32)
        q = p.pooled(node_map)                               # p owns [0,4) and [4,8): 3 kept below row 4, 1 above
        ok &= q.bounds == [0, 3, 4] and q.n == 4
        q2 = p2.pooled(node_map)                             # [0,5) keeps 3, [5,8) keeps 1
        ok &= q2.bounds == [0, 3, 4]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_gather_rows_and_pooled_bounds_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_thread_comm_matches_concatenation():
    import threading
    comms = ThreadComm.make(3)
    full = torch.arange(30, dtype=torch.float32).view(10, 3)
    bounds = [0, 4, 4, 10]                                   # the middle rank owns no rows
    out = [None] * 3

    def run(r):
        p = Part(comms[r], bounds)
        out[r] = p.gather(p.local(full).clone())

    ts = [threading.Thread(target=run, args=(r,)) for r in range(3)]
    [t.start() for t in ts]
    [t.join(30) for t in ts]
    assert all(o is not None and torch.equal(o, full) for o in out)
