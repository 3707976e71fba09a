"""Host -> device staging for serving loops: the reference moves one graph at a time with `graph.to(device)` right before
the forward (evaluation/predictor.py:205), which leaves the 307 MB feature copy of a 100k-node slide fully exposed.
DevicePrefetcher copies graph i+1 from pinned host memory on a side stream while graph i is being processed."""
import torch

from .graph import Data


class DevicePrefetcher:
    """Iterate over host graphs (Data with pinned tensors) and yield device copies, one graph ahead.

        for d in DevicePrefetcher(host_graphs, device): out = model(d)
    """

    def __init__(self, graphs, device, rows=None):
        """rows: optional (a, b) — copy only these rows of x (the rows a rank owns in node-partitioned execution)."""
        self.graphs, self.device, self.rows = graphs, torch.device(device), rows
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [{}, {}]          # two sets of device staging tensors, reused (no allocator traffic in the loop)
        self._n = 0

    def _buf(self, slot, key, host):
        t = slot.get(key)
        if t is None or t.shape != host.shape or t.dtype != host.dtype:
            t = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            slot[key] = t
        return t

    def _stage(self, g):
        slot = self._slots[self._n & 1]
        self._n += 1
        # the slot was last used two graphs ago: everything enqueued on the consumer's stream so far includes that forward
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            x = g.x if self.rows is None else g.x[self.rows[0]:self.rows[1]]
            kw = {}
            for k in ("edge_index", "edge_attr", "pos", "batch"):
                v = getattr(g, k, None)
                kw[k] = None if v is None else self._buf(slot, k, v).copy_(v, non_blocking=True)
            batch = kw.pop("batch")
            d = Data(x=self._buf(slot, "x", x).copy_(x, non_blocking=True), **kw)
            if batch is not None:
                d.batch = batch
            if self.rows is not None:
                d.num_nodes = int(g.x.size(0))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return d, ev

    def __iter__(self):
        it = iter(self.graphs)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            d, ev = nxt
            torch.cuda.current_stream(self.device).wait_event(ev)
            try:
                nxt = self._stage(next(it))          # overlaps with the consumer's forward on graph `d`
            except StopIteration:
                nxt = None
            yield d
