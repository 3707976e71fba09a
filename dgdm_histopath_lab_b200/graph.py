"""Graph containers (duck-typed stand-ins for torch_geometric.data.Data / Batch, which the reference imports at
models/dgdm_model.py:11) and the per-graph structure cache used by the kernels.

If torch_geometric is installed its Data/Batch objects are accepted as well (isinstance check in
DGDMModel._validate_forward_inputs mirrors models/dgdm_model.py:650)."""
import copy

import torch

from . import ops

try:  # optional — not installed in the build image
    from torch_geometric.data import Data as _PygData, Batch as _PygBatch   # type: ignore
    _PYG = (_PygData, _PygBatch)
except Exception:  # pragma: no cover
    _PYG = ()


class Data:
    """Attribute bag: x [N,F], edge_index [2,E] int64, edge_attr [E,D]|None, pos [N,2]|None, batch [N]|None, y."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos = x, edge_index, edge_attr, y, pos
        self.batch = None
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        n = self.__dict__.get("_num_nodes")
        if n is not None:
            return n
        return int(self.x.size(0)) if self.x is not None else 0

    @num_nodes.setter
    def num_nodes(self, v):
        self.__dict__["_num_nodes"] = v

    @property
    def num_edges(self):
        return int(self.edge_index.size(1)) if self.edge_index is not None else 0

    def _map(self, fn):
        out = copy.copy(self)
        out.__dict__ = dict(self.__dict__)
        out.__dict__.pop("_dgdm_ctx", None)
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                out.__dict__[k] = fn(v)
        return out

    def clone(self):
        return self._map(lambda t: t.clone())

    def to(self, device, non_blocking=False):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        """Concatenate graphs: x/pos/edge_attr stacked, edge_index offset, `batch` = graph id per node."""
        xs, eis, eas, ps, bs, off = [], [], [], [], [], 0
        for i, d in enumerate(data_list):
            n = d.x.size(0)
            xs.append(d.x)
            eis.append(d.edge_index + off)
            if d.edge_attr is not None:
                eas.append(d.edge_attr)
            if d.pos is not None:
                ps.append(d.pos)
            bs.append(torch.full((n,), i, dtype=torch.long, device=d.x.device))
            off += n
        out = cls(x=torch.cat(xs), edge_index=torch.cat(eis, dim=1),
                  edge_attr=torch.cat(eas) if len(eas) == len(data_list) and eas else None,
                  pos=torch.cat(ps) if len(ps) == len(data_list) and ps else None)
        out.batch = torch.cat(bs)
        out.num_graphs = len(data_list)
        return out


GRAPH_TYPES = (Data, Batch) + _PYG


class GraphContext:
    """Everything that depends only on (edge_index, batch) and is reused by all 24 convolutions of a forward:
    CSR by destination, GCN norm, segment pointers (SURVEY §8 identity (ii)). Cached on the data object."""

    def __init__(self, edge_index, batch, num_nodes):
        self.N = int(num_nodes)
        self.csr = ops.csr_build(edge_index, self.N)
        dev = edge_index.device
        if batch is None:
            self.num_graphs = 1
            self.seg_host = [0, self.N]
            self.sorted = True
            self.seg_ptr = torch.arange(2, dtype=torch.int32, device=dev) * self.N      # device-side: capturable into a CUDA graph
        else:
            counts = torch.bincount(batch).cpu()          # one host sync per graph object (the reference does .item() per call)
            self.num_graphs = int(counts.numel())
            self.seg_host = [0] + torch.cumsum(counts, 0).tolist()
            self.sorted = bool((batch[1:] >= batch[:-1]).all().item()) if batch.numel() > 1 else True
            self.seg_ptr = torch.tensor(self.seg_host, dtype=torch.int32, device=dev)
        self.agg_attr = None
        self._key = None


def get_context(data, num_nodes=None):
    """Build (or fetch) the GraphContext of a data object; keyed on the identity/version of its index tensors.
    num_nodes: N of the whole batch when data.x holds only a rank's rows (node-partitioned execution)."""
    ei, batch = data.edge_index, getattr(data, "batch", None)
    n = int(data.x.size(0)) if num_nodes is None else int(num_nodes)
    key = (ei.data_ptr(), ei._version, tuple(ei.shape),
           None if batch is None else (batch.data_ptr(), batch._version, batch.numel()), n)
    ctx = data.__dict__.get("_dgdm_ctx") if hasattr(data, "__dict__") else None
    if ctx is None or ctx._key != key:
        ctx = GraphContext(ei, batch, n)
        ctx._key = key
        try:
            data.__dict__["_dgdm_ctx"] = ctx
        except Exception:
            pass
    return ctx
