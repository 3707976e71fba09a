"""Multi-GPU execution of the DGDM forward: one process per GPU, 1-D node partition.

The reference has no multi-GPU forward at all (its only collective is Lightning DDP's gradient all-reduce,
cli/train.py:346-350), so this module is new design, not a mirror:

  * the GRAPH STRUCTURE (CSR, norm, pooled CSRs, top-k selections) is replicated: it is integer work that costs
    < 2 % of a forward and replicating it removes every index exchange;
  * every FEATURE matrix is partitioned by contiguous node ranges: rank r owns rows [bounds[r], bounds[r+1]) of every
    activation. Row-wise work (GEMMs, LayerNorm, attention query rows, un-pool) runs on the owned rows only;
  * each GraphConvolution.propagate needs the post-node_lin rows of the sources of its in-edges: the "halo". On the
    BASELINE tile graph half of the links are uniformly random, so the halo of a rank is (1 - 1/P) of the graph
    and the exchange is an all-gather of the [N, C] matrix (NCCL over NVLink; C is the narrower, post-projection
    width). The accumulation order inside a destination row is the single-GPU order, so partitioned results are
    bitwise identical to the single-GPU ones;
  * spatial attention is context parallelism: queries = owned rows, keys/values = all-gathered q|k|v rows;
  * AdaptiveGraphPooling all-gathers the N scores; every rank runs the same exact top-k and keeps the selected nodes
    it already owns (perm is ascending, so ownership stays contiguous: new bounds = selected-count prefix sums);
  * GlobalAttentionPool merges per-rank (max, sum, weighted value) partials after one small all-gather.

`Comm` abstracts the collective so that the bookkeeping is testable on CPU (gloo, world_size 2) and the whole
partitioned forward is testable on ONE GPU with in-process ranks (`ThreadComm`, threads + barriers).
"""
import threading

import torch


def even_bounds(n, world):
    """Contiguous row ranges of ceil(n/world) rows (the last ranks may be short or empty)."""
    rows = -(-n // world)
    return [min(n, r * rows) for r in range(world + 1)]


class Comm:
    """all_gather_rows(t, bounds): concatenate the ranks' row blocks (rank r contributes bounds[r+1]-bounds[r] rows)."""
    rank = 0
    world = 1

    def all_gather_rows(self, t, bounds):
        raise NotImplementedError

    def barrier(self):
        pass


class TorchDistComm(Comm):
    """torch.distributed (NCCL on GPUs, gloo on CPU for tests). Even blocks -> one all_gather_into_tensor straight into the
    result; uneven blocks (pooled U-Net levels) -> padded all-gather + one compaction copy."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather_rows(self, t, bounds):
        dist = self._dist
        sizes = [bounds[r + 1] - bounds[r] for r in range(self.world)]
        assert t.size(0) == sizes[self.rank], (t.shape, sizes, self.rank)
        tail = tuple(t.shape[1:])
        mx = max(sizes)
        if all(s == mx for s in sizes):
            out = torch.empty((mx * self.world,) + tail, dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
            return out
        pad = torch.zeros((mx,) + tail, dtype=t.dtype, device=t.device)
        pad[: t.size(0)] = t
        buf = torch.empty((mx * self.world,) + tail, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad, group=self.group)
        return torch.cat([buf[r * mx: r * mx + sizes[r]] for r in range(self.world)], dim=0)

    def barrier(self):
        self._dist.barrier(group=self.group)


class SymmComm(TorchDistComm):
    """TorchDistComm + the fused halo exchange: the projected rows of a GraphConvolution are written by the GEMM epilogue
    itself into every rank's copy of the [N, C] matrix — P2P stores over NVLink into symmetric memory, or one NVLS multicast
    store per tile when the switch supports it — so the all-gather costs no extra kernel and overlaps the GEMM tile by tile.
    Two symmetric buffers alternate between consecutive convolutions; a device-side barrier (signal pads, on the compute
    stream) separates the stores of convolution k from the SpMM that reads them, and — because every rank passes barrier
    k+1 only after its own SpMM k — also protects buffer k%2 from being overwritten by convolution k+2 while still in use."""

    def __init__(self, max_floats, device, group=None, use_multicast=True):
        super().__init__(group)
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self._pg = group if group is not None else dist.group.WORLD
        self.max_floats = int(max_floats)
        self.bufs, self.hdls = [], []
        for _ in range(2):
            t = symm_mem.empty(self.max_floats, dtype=torch.float32, device=device)
            self.hdls.append(symm_mem.rendezvous(t, self._pg))
            self.bufs.append(t)
        # measured on B200 (tools/halo_bench.py, [100k, 512] fp32): 4 GPUs — P2P stores 275 us vs NVLS multicast 326 us vs NCCL
        # all_gather 320 us (+86 us GEMM); 8 GPUs — 319 vs 326 vs 316 (+45) us: all NVLink-inbound-bound (~560 GB/s per GPU).
        # Multicast needs 1/(P-1) of the SM->NVLink egress, so it is used from 8 ranks up, plain P2P stores below.
        self.multicast = bool(use_multicast) and self.world >= 8 and all(int(getattr(h, "multicast_ptr", 0) or 0) != 0 for h in self.hdls)
        self._k = 0

    def fused_linear_gather(self, linear_fn, part, n_cols):
        """linear_fn(out, peers, multicast) runs the GEMM whose epilogue writes this rank's rows [a, b) of the [n, n_cols]
        matrix into `out` (a view of the local symmetric buffer) and into the peers' buffers. Returns the full matrix."""
        need = part.n * n_cols
        if need > self.max_floats:
            return None
        i = self._k & 1
        self._k += 1
        buf, hdl = self.bufs[i], self.hdls[i]
        full = buf[:need].view(part.n, n_cols)
        out = full[part.a:part.b]
        off = part.a * n_cols * 4
        if self.multicast:
            peers, mc = [int(hdl.multicast_ptr) + off], True
        else:
            peers, mc = [int(hdl.buffer_ptrs[r]) + off for r in range(self.world) if r != self.rank], False
        if part.n_local > 0:
            linear_fn(out, peers, mc)
        hdl.barrier(channel=0)
        return full


class ThreadComm(Comm):
    """In-process ranks (one Python thread each, one shared GPU) for tests: the exchange is a device synchronise + a
    threading.Barrier around a shared slot list. Create with ThreadComm.make(world) -> list of per-rank comms."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self._sh = rank, world, shared

    @classmethod
    def make(cls, world):
        shared = {"slots": [None] * world, "bar": threading.Barrier(world)}
        return [cls(r, world, shared) for r in range(world)]

    def all_gather_rows(self, t, bounds):
        sh = self._sh
        if t.is_cuda:
            torch.cuda.synchronize()
        sh["slots"][self.rank] = t
        sh["bar"].wait()
        out = torch.cat([sh["slots"][r][: bounds[r + 1] - bounds[r]] for r in range(self.world)], dim=0)
        if t.is_cuda:
            torch.cuda.synchronize()
        sh["bar"].wait()
        return out

    def barrier(self):
        self._sh["bar"].wait()


class Part:
    """Partition of one node set (one U-Net level): comm + row bounds. `a`, `b` = the rows this rank owns."""

    def __init__(self, comm, bounds):
        self.comm, self.bounds = comm, [int(v) for v in bounds]
        self.a, self.b = self.bounds[comm.rank], self.bounds[comm.rank + 1]
        self.n = self.bounds[-1]

    @property
    def rows(self):
        return (self.a, self.b)

    @property
    def n_local(self):
        return self.b - self.a

    def gather(self, t):
        """[n_local, ...] -> [n, ...] (all ranks' blocks in node order)."""
        return self.comm.all_gather_rows(t, self.bounds)

    def local(self, full):
        return full[self.a:self.b]

    def pooled(self, node_map):
        """Partition of the pooled node set: a selected node stays on its owner. node_map [n] int32 (-1 = dropped) is
        replicated, so every rank derives the same bounds (one small host read per pooling level)."""
        sel = (node_map >= 0).to(torch.int32)
        csum = torch.cumsum(sel, 0)
        idx = torch.tensor([v - 1 for v in self.bounds[1:]], dtype=torch.long, device=node_map.device).clamp_(min=0)
        ends = csum[idx].tolist() if self.n > 0 else [0] * self.comm.world
        nb = [0]
        for r in range(self.comm.world):
            nb.append(int(ends[r]) if self.bounds[r + 1] > 0 else 0)
        return Part(self.comm, nb)
