"""Graph files and batching for the serving loop (reference: data/dataset.py:190-240 `_load_graph` for
.pt / .pkl / .h5, and the PyG DataLoader collate the trainer relies on).

Besides the reference's formats this module defines `.dgdm`, a flat binary container whose arrays can be memory-mapped and
copied to the device without parsing: fp32 features / positions / edge attributes, the int64 `edge_index` the reference API
exposes, and — optionally — the CSR-by-destination (int32 rowptr / col / eid, original edge order kept inside each row) that
every forward otherwise rebuilds with a device sort. `load_graph(..., device=...)` attaches a ready GraphContext in that case.
"""
import json
import os
import pickle
import struct

import numpy as np
import torch

from . import ops
from .graph import Batch, Data, GraphContext

MAGIC = b"DGDMGRF1"
_DTYPES = {"float32": np.float32, "float64": np.float64, "int64": np.int64, "int32": np.int32}


def csr_by_destination(edge_index, num_nodes):
    """Host-side CSR by destination with the original edge order inside each row (stable sort) — the same structure
    dgdm_csr_build produces on the device. Returns int32 (rowptr [N+1], col [E], eid [E])."""
    ei = np.asarray(edge_index)
    dst = ei[1]
    eid = np.argsort(dst, kind="stable").astype(np.int32)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, ei[0][eid].astype(np.int32), eid


def save_graph(data, path, with_csr=True):
    """Write `data` (x, edge_index, optional pos / edge_attr / y / batch) as a .dgdm file."""
    arrays = {}
    for k in ("x", "edge_index", "pos", "edge_attr", "y", "batch", "edge_type"):
        v = getattr(data, k, None)
        if v is not None:
            arrays[k] = np.ascontiguousarray(v.detach().cpu().numpy())
    n = int(arrays["x"].shape[0])
    if with_csr:
        arrays["csr_rowptr"], arrays["csr_col"], arrays["csr_eid"] = csr_by_destination(arrays["edge_index"], n)
    meta, blobs, off = {"num_nodes": n, "arrays": {}}, [], 0
    for k, a in arrays.items():
        if str(a.dtype) not in _DTYPES:
            a = a.astype(np.float32 if a.dtype.kind == "f" else np.int64)
        pad = (-off) % 64
        off += pad
        blobs.append((pad, a))
        meta["arrays"][k] = {"dtype": str(a.dtype), "shape": list(a.shape), "offset": off}
        off += a.nbytes
    for k in ("slide_id", "num_patches"):
        if hasattr(data, k) and isinstance(getattr(data, k), (str, int, float)):
            meta[k] = getattr(data, k)
    head = json.dumps(meta).encode()
    head += b" " * ((-(len(MAGIC) + 8 + len(head))) % 64)
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<Q", len(head)) + head)
        for pad, a in blobs:
            f.write(b"\0" * pad)
            f.write(a.tobytes())
    return path


def _load_dgdm(path, device=None, pin=False):
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a .dgdm graph file")
        (hl,) = struct.unpack("<Q", f.read(8))
        meta = json.loads(f.read(hl).decode())
        base = 16 + hl
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    t = {}
    for k, d in meta["arrays"].items():
        cnt = int(np.prod(d["shape"])) if d["shape"] else 1
        a = np.frombuffer(mm, dtype=_DTYPES[d["dtype"]], count=cnt, offset=base + d["offset"]).reshape(d["shape"])
        ten = torch.from_numpy(np.array(a))                     # one copy out of the page cache
        if pin and device is None:
            ten = ten.pin_memory()
        t[k] = ten if device is None else ten.to(device, non_blocking=True)
    data = Data(x=t["x"], edge_index=t["edge_index"], edge_attr=t.get("edge_attr"), y=t.get("y"), pos=t.get("pos"))
    if "batch" in t:
        data.batch = t["batch"]
    if "edge_type" in t:
        data.edge_type = t["edge_type"]
    for k in ("slide_id", "num_patches"):
        if k in meta:
            setattr(data, k, meta[k])
    if "csr_rowptr" in t and device is not None and torch.device(device).type == "cuda":
        attach_csr(data, t["csr_rowptr"], t["csr_col"], t["csr_eid"])
    elif "csr_rowptr" in t:
        data.csr_host = (t["csr_rowptr"], t["csr_col"], t["csr_eid"])
    return data


def attach_csr(data, rowptr, col, eid):
    """Install a GraphContext built from a stored CSR (device int32 tensors) so that forward() skips dgdm_csr_build. The GCN
    normalisation is derived with the kernel's own arithmetic: dinv = 1/sqrt(in-degree + 1), norm[e] = dinv[src] * dinv[dst]
    (core/graph_layers.py:81-84) — IEEE sqrt / divide / multiply, bitwise equal to dgdm_csr_build's output."""
    n = int(rowptr.numel() - 1)
    e = int(col.numel())
    deg = (rowptr[1:] - rowptr[:-1] + 1).to(torch.float32)
    dinv = 1.0 / torch.sqrt(deg)
    dst = torch.repeat_interleave(torch.arange(n, device=rowptr.device), (rowptr[1:] - rowptr[:-1]).long())
    norm = dinv[col.long()] * dinv[dst]
    ctx = GraphContext.__new__(GraphContext)
    ctx.N = n
    ctx.csr = ops.CSR(rowptr.contiguous(), col.contiguous(), eid.contiguous(), norm.contiguous(), dinv.contiguous(), n, e)
    batch = getattr(data, "batch", None)
    if batch is None:
        ctx.num_graphs, ctx.seg_host, ctx.sorted = 1, [0, n], True
    else:
        counts = torch.bincount(batch).cpu()
        ctx.num_graphs = int(counts.numel())
        ctx.seg_host = [0] + torch.cumsum(counts, 0).tolist()
        ctx.sorted = bool((batch[1:] >= batch[:-1]).all().item()) if batch.numel() > 1 else True
    ctx.seg_ptr = torch.tensor(ctx.seg_host, dtype=torch.int32, device=rowptr.device)
    ctx.agg_attr = None
    ei = data.edge_index
    ctx._key = (ei.data_ptr(), ei._version, tuple(ei.shape), None if batch is None else (batch.data_ptr(), batch._version, batch.numel()), n)
    data.__dict__["_dgdm_ctx"] = ctx
    return data


def load_graph(path, device=None, pin=False):
    """data/dataset.py:190-200 `_load_graph`: .pt (torch.load), .pkl (pickle), .h5 (h5py: node_features, edge_index, edge_attr,
    node_pos, labels — :203-240), plus .dgdm. Like the reference, .pt / .pkl files are UNPICKLED (arbitrary code execution):
    load them from trusted sources only; .dgdm and .h5 are plain arrays."""
    path = str(path)
    ext = os.path.splitext(path)[1]
    if ext == ".dgdm":
        return _load_dgdm(path, device, pin)
    if ext == ".pt":
        obj = torch.load(path, weights_only=False)
    elif ext == ".pkl":
        with open(path, "rb") as f:
            obj = pickle.load(f)
    elif ext == ".h5":
        try:
            import h5py
        except ImportError as e:
            raise ImportError("reading .h5 graph files needs h5py (not installed in this image); convert with save_graph()") from e
        with h5py.File(path, "r") as f:
            obj = Data(x=torch.tensor(f["node_features"][:], dtype=torch.float), edge_index=torch.tensor(f["edge_index"][:], dtype=torch.long),
                       edge_attr=torch.tensor(f["edge_attr"][:], dtype=torch.float) if "edge_attr" in f else None,
                       pos=torch.tensor(f["node_pos"][:], dtype=torch.float) if "node_pos" in f else None,
                       y=torch.tensor(f["labels"][:], dtype=torch.long) if "labels" in f else None)
            if "metadata" in f:
                for key in f["metadata"].attrs:
                    setattr(obj, key, f["metadata"].attrs[key])
    else:
        raise ValueError(f"Unsupported graph format: {ext}")
    if isinstance(obj, dict):
        obj = Data(**obj)
    return obj if device is None else obj.to(device)


def collate_to_device(graphs, device, stream=None):
    """On-device batching of host graphs (the PyG DataLoader collate, done with one H2D copy per tensor into a preallocated batch):
    x / pos / edge_attr rows are concatenated, edge_index is offset on the device, `batch` holds the graph id per node."""
    device = torch.device(device)
    ns = [int(g.x.size(0)) for g in graphs]
    es = [int(g.edge_index.size(1)) for g in graphs]
    N, E = sum(ns), sum(es)
    f = graphs[0].x.size(1)
    has_pos = all(getattr(g, "pos", None) is not None for g in graphs)
    has_ea = all(getattr(g, "edge_attr", None) is not None and g.edge_attr.size(0) == g.edge_index.size(1) for g in graphs)
    ea_w = max((g.edge_attr.size(1) for g in graphs), default=0) if has_ea else 0
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream(device))
    with ctx:
        x = torch.empty(N, f, dtype=torch.float32, device=device)
        ei = torch.empty(2, E, dtype=torch.int64, device=device)
        pos = torch.empty(N, 2, dtype=torch.float32, device=device) if has_pos else None
        ea = torch.zeros(E, ea_w, dtype=torch.float32, device=device) if has_ea else None
        batch = torch.empty(N, dtype=torch.int64, device=device)
        no, eo = 0, 0
        for i, g in enumerate(graphs):
            n, e = ns[i], es[i]
            x[no:no + n].copy_(g.x, non_blocking=True)
            if e:
                ei[:, eo:eo + e].copy_(g.edge_index, non_blocking=True)
                ei[:, eo:eo + e] += no
            if has_pos:
                pos[no:no + n].copy_(g.pos, non_blocking=True)
            if has_ea and e:
                ea[eo:eo + e, :g.edge_attr.size(1)].copy_(g.edge_attr, non_blocking=True)
            batch[no:no + n] = i
            no, eo = no + n, eo + e
    out = Batch(x=x, edge_index=ei, edge_attr=ea, pos=pos)
    out.batch = batch
    out.num_graphs = len(graphs)
    return out


class GraphFileDataset:
    """Minimal mirror of the graph branch of data/dataset.py's SlideDataset: a directory (or list) of graph files."""

    EXTS = (".dgdm", ".pt", ".pkl", ".h5")

    def __init__(self, root_or_files, device=None, pin=False):
        if isinstance(root_or_files, (str, os.PathLike)):
            root = str(root_or_files)
            self.files = sorted(os.path.join(root, f) for f in os.listdir(root) if f.endswith(self.EXTS))
        else:
            self.files = [str(f) for f in root_or_files]
        self.device, self.pin = device, pin

    def __len__(self):
        return len(self.files)

    def __getitem__(self, idx):
        return load_graph(self.files[idx], self.device, self.pin)
