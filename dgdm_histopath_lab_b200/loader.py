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
        slot_index = self._n & 1
        slot = self._slots[slot_index]
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
            d.__dict__["_stage_slot"] = slot_index           # which of the two staging buffer sets backs this graph (GraphServer)
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


class _Deferred:
    """What `_finish` needs of a forward whose validation read-back is still pending."""

    def __init__(self, validation_buffer, n_x):
        self.validation_buffer, self.n_x = validation_buffer, n_x


class GraphServer:
    """Serving loop for a stream of host graphs (the reference's `evaluation/predictor.py:188-215` loop, one slide graph at a time):
    DevicePrefetcher staging + ONE captured CUDA graph per staging slot + a read-back stream. While graph i runs, graph i+1 is
    copied from pinned host memory on a side stream into the other slot; a slot whose tensors keep their shapes replays its
    captured forward (CSR build and input validation included) instead of issuing ~150 launches. The replay of graph i+1 is
    enqueued BEFORE the results of graph i are read back (the two slots own separate output buffers), and the read-back — the
    fetched outputs plus the 32 validation bytes — runs on its own stream behind an event, so the device never waits for the
    host. Validation keeps the forward's error behaviour (NaN / Inf features, out-of-range edge indices raise
    `ModelInferenceError` at the position of the offending graph). Graphs of a new shape re-capture their slot; node-partitioned
    models (no graph capture) and batched inputs run eager launches through the same pipeline (validation kernel enqueued without
    its read-back, results read back behind the next graph's forward).

        server = GraphServer(model, device, return_embeddings=True)
        for out in server.run(host_graphs):          # fetched keys are host tensors; the rest are the slot's static device buffers,
            emb = out["graph_embedding"]             # valid until the next-but-one graph: copy what must live longer
    """

    def __init__(self, model, device, mode="inference", return_embeddings=False, fetch=("graph_embedding",), use_graphs=True):
        self.model, self.device, self.mode, self.return_embeddings = model, torch.device(device), mode, return_embeddings
        self.fetch, self.use_graphs = tuple(fetch), use_graphs
        self._runs = {}                      # slot index -> (key of the slot's tensors, captured callable)
        self._pre = DevicePrefetcher([], self.device)        # ONE prefetcher for the server's lifetime: the captured graphs read its buffers
        self._rb_stream = torch.cuda.Stream(device=self.device)
        self._host = [{}, {}]                # pinned read-back buffers, two sets (results of consecutive graphs)
        self._n_out = 0
        self.replays, self.captures, self.eager = 0, 0, 0

    @staticmethod
    def _key(d):
        return tuple((k, tuple(getattr(d, k).shape), getattr(d, k).data_ptr()) for k in ("x", "edge_index", "edge_attr", "pos")
                     if getattr(d, k, None) is not None)

    def _pinned(self, hs, key, t):
        b = hs.get(key)
        if b is None or b.shape != t.shape or b.dtype != t.dtype:
            b = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            hs[key] = b
        return b

    def _finish(self, pending):
        """Read the fetched outputs and the validation bytes of a replay back on the read-back stream, check, build the result."""
        out, run, ev = pending
        hs = self._host[self._n_out & 1]
        self._n_out += 1
        res = dict(out)
        with torch.cuda.stream(self._rb_stream):
            self._rb_stream.wait_event(ev)
            for k in self.fetch:
                v = out.get(k)
                if torch.is_tensor(v):
                    res[k] = self._pinned(hs, k, v).copy_(v, non_blocking=True)
            vb = run.validation_buffer
            flags = None if vb is None else self._pinned(hs, "_validation", vb).copy_(vb, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._rb_stream)
        done.synchronize()
        if flags is not None:
            from .errors import ModelInferenceError
            try:
                self.model._check_validation_result(flags[:4].tolist(), flags[4:].view(torch.float32).tolist(), run.n_x)
            except Exception as e:
                raise ModelInferenceError(f"Input validation failed: {e}")
        return res

    def run(self, host_graphs, rows=None):
        self._pre.graphs, self._pre.rows = host_graphs, rows
        pending = None
        for d in self._pre:
            slot = d.__dict__.get("_stage_slot", 0)
            if not self.use_graphs or getattr(self.model, "_comm", None) is not None or getattr(d, "batch", None) is not None:
                # eager launches (node-partitioned model: NCCL and symmetric-memory barriers are not capturable; batches: a host
                # read for the segment sizes), same pipeline: the validation kernel is enqueued without its read-back, so the host
                # enqueues the whole forward without synchronising, and the results are read back behind the next graph's forward
                from .errors import ModelInferenceError
                try:
                    vbuf = self.model._validate_forward_inputs(d, self.mode, False, self.return_embeddings, deferred=True)
                except Exception as e:
                    raise ModelInferenceError(f"Input validation failed: {e}")
                out = self.model(d, self.mode, False, self.return_embeddings, validate=False)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self.eager += 1
                if pending is not None:
                    yield self._finish(pending)
                pending = (out, _Deferred(vbuf, int(d.x.numel())), ev)
                continue
            key = self._key(d)
            ent = self._runs.get(slot)
            if ent is None or ent[0] != key:
                # first graph of this shape in this slot: validated eager warm-up, then capture (the slot's tensors are static)
                ent = (key, self.model.graphed(d, self.mode, self.return_embeddings, warmup=1, rebuild_structure=True, validate=True))
                self._runs[slot] = ent
                self.captures += 1
            out = ent[1]()                   # enqueue the replay of THIS graph ...
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.replays += 1
            if pending is not None:          # ... and only then read the previous graph's results back
                yield self._finish(pending)
            pending = (out, ent[1], ev)
        if pending is not None:
            yield self._finish(pending)
