"""TissueGraphBuilder — drop-in for preprocessing/tissue_graph_builder.py:42-447 (build_graph path).

The O(N^2) parts (768-d Gram / Euclidean kNN / cosine, 2-D kNN) run on the GPU through the C-ABI; candidate
filtering, duplicate removal and edge emission reproduce the reference's order exactly with vectorised numpy
(the reference uses Python loops over dataclass objects; the result — indices AND order — is the same).
"""
import logging
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .graph import Data
from .ops import KernelError, _check, _p, _stream


@dataclass
class PatchInfo:
    """preprocessing/slide_processor.py:34-43."""
    x: int
    y: int
    level: int
    magnification: float
    patch_id: str
    tissue_percentage: float
    features: Optional[np.ndarray] = None


@dataclass
class SlideData:
    """preprocessing/slide_processor.py:46-53."""
    slide_id: str
    patches: List[PatchInfo]
    metadata: Dict
    thumbnail: Optional[np.ndarray] = None
    tissue_mask: Optional[np.ndarray] = None


@dataclass
class GraphNode:
    node_id: str
    patch_info: PatchInfo
    features: np.ndarray
    spatial_coords: Tuple[float, float]
    tissue_type: Optional[str] = None
    confidence: float = 1.0


@dataclass
class GraphEdge:
    source_id: str
    target_id: str
    edge_type: str
    weight: float
    features: Optional[np.ndarray] = None


TC_MIN_ROWS = 4096          # below this the exact fp64 CUDA-core path is used directly
KNN_STATS = {"rows": 0, "flagged": 0}


def knn_cosine(features, k, q0=0, nq=None, chunk=None, mode=None):
    """features: float32 CUDA [n,d]. Returns (idx int32 [nq,k], cos float32 [nq,k], dist float64 [nq,k]) for rows q0..q0+nq
    (q0/nq = the query rows one rank owns when the rows are sharded across GPUs).
    mode 'exact': fp64 Gram on CUDA cores. mode 'tc' (default for n >= TC_MIN_ROWS): tcgen05 candidate pass + exact fp64
    re-rank + exact recomputation of the rows whose top-k could not be proved (KNN_STATS counts them); same results."""
    n, d = features.shape
    nq = n - q0 if nq is None else nq
    if mode is None:
        mode = "tc" if (n >= TC_MIN_ROWS and d % 4 == 0 and k <= 32) else "exact"
    if mode == "tc":
        return _knn_cosine_tc(features, k, q0, nq, chunk)
    return _knn_cosine_exact(features, k, q0, nq, chunk)


def _knn_cosine_exact(features, k, q0, nq, chunk=None):
    lib = _lib.load()
    n, d = features.shape
    dev = features.device
    idx = torch.empty(nq, k, dtype=torch.int32, device=dev)
    cos = torch.empty(nq, k, dtype=torch.float32, device=dev)
    dist = torch.empty(nq, k, dtype=torch.float64, device=dev)
    if chunk is None:
        chunk = max(64, min(nq, int(2e9 // (8 * n))))          # <= 2 GB of fp64 distance rows per call
    for c0 in range(0, nq, chunk):
        cn = min(chunk, nq - c0)
        wb = int(lib.dgdm_knn_cosine_work_bytes(cn, n, d, k))
        work = torch.empty(wb, dtype=torch.uint8, device=dev)
        _check(lib.dgdm_knn_cosine(_p(features), n, d, q0 + c0, cn, k, _p(idx[c0:]), _p(cos[c0:]), _p(dist[c0:]), _p(work), wb,
                                   _stream()), "dgdm_knn_cosine")
    return idx, cos, dist


def _knn_cosine_tc(features, k, q0, nq, chunk=None):
    from . import ops
    lib = _lib.load()
    n, d = features.shape
    dev = features.device
    # power-of-two prescale so that the fp16 hi/lo planes of the largest entries sit well inside the normal fp16 range; it is
    # exact in fp32/fp64, so rankings, ties and cosines are unchanged and distances are rescaled exactly at the end
    amax = float(features.abs().max().item())
    scale = 2.0 ** round(math.log2(8.0 / amax)) if amax > 0 else 1.0
    n_pad = (n + 15) // 16 * 16
    fs = torch.zeros(n_pad, d, dtype=torch.float32, device=dev)
    torch.mul(features, scale, out=fs[:n])
    img = ops.SplitWeight(fs)
    sq = torch.empty(n_pad, dtype=torch.float64, device=dev)
    _check(lib.dgdm_row_sqnorm_f64(_p(fs), n_pad, d, _p(sq), _stream()), "dgdm_row_sqnorm_f64")
    idx = torch.empty(nq, k, dtype=torch.int32, device=dev)
    cos = torch.empty(nq, k, dtype=torch.float32, device=dev)
    dist = torch.empty(nq, k, dtype=torch.float64, device=dev)
    flags = torch.empty(nq, dtype=torch.int32, device=dev)
    chunk = 16384 if chunk is None else chunk
    wb = int(lib.dgdm_knn_cosine_tc_work_bytes(min(chunk, nq)))
    work = torch.empty(wb, dtype=torch.uint8, device=dev)
    for c0 in range(0, nq, chunk):
        cn = min(chunk, nq - c0)
        _check(lib.dgdm_knn_cosine_tc(_p(fs), img.hi_ptr, _p(sq), n, n_pad, d, q0 + c0, cn, k, _p(idx[c0:]), _p(cos[c0:]), _p(dist[c0:]),
                                      _p(flags[c0:]), _p(work), wb, _stream()), "dgdm_knn_cosine_tc")
    bad = torch.nonzero(flags).reshape(-1).tolist()
    KNN_STATS["rows"] += nq
    KNN_STATS["flagged"] += len(bad)
    if bad:       # rows whose top-k could not be proved from 32 candidates: exact fp64 path, one row at a time
        fv = fs[:n]
        for r in bad:
            i1, c1, d1 = _knn_cosine_exact(fv, k, q0 + r, 1)
            idx[r], cos[r], dist[r] = i1[0], c1[0], d1[0]
    dist /= scale
    return idx, cos, dist


GRID_MIN_ROWS = 2048        # below this the brute-force 2-D kNN is used


def knn2d(coords, k, q0=0, nq=None, chunk=None, mode=None):
    """coords: float64 CUDA [n,2]. Returns (idx int32 [nq,k], dist float64 [nq,k]). mode 'grid' (default for n >= GRID_MIN_ROWS):
    uniform-grid ring search, O(N); mode 'brute': fp64 distance rows, O(N^2). Identical results."""
    lib = _lib.load()
    n = coords.size(0)
    nq = n - q0 if nq is None else nq
    dev = coords.device
    idx = torch.empty(nq, k, dtype=torch.int32, device=dev)
    dist = torch.empty(nq, k, dtype=torch.float64, device=dev)
    if mode is None:
        mode = "grid" if n >= GRID_MIN_ROWS else "brute"
    if mode == "grid":
        work = torch.empty(int(lib.dgdm_knn2d_grid_work_bytes(n)), dtype=torch.uint8, device=dev)
        _check(lib.dgdm_knn2d(_p(coords), n, q0, nq, k, _p(idx), _p(dist), 1, _p(work), _stream()), "dgdm_knn2d")
        return idx, dist
    if chunk is None:
        chunk = max(64, min(nq, int(2e9 // (8 * n))))
    for c0 in range(0, nq, chunk):
        cn = min(chunk, nq - c0)
        _check(lib.dgdm_knn2d(_p(coords), n, q0 + c0, cn, k, _p(idx[c0:]), _p(dist[c0:]), 0, None, _stream()), "dgdm_knn2d")
    return idx, dist


def candidate_edges(sp_idx, sp_dist, mo_idx, mo_cos, threshold, row0=0):
    """Reference loops :302-318 and :337-350 vectorised. Inputs are numpy [nq,k+1] arrays for query rows row0..;
    column 0 (self) is skipped. Returns dict of flat arrays in the reference's emission order (spatial then morphological,
    row-major): src, tgt, type(0/1), weight float64, f0, f1 (edge features zero-padded to width 2, P7)."""
    out = []
    if sp_idx is not None and sp_idx.shape[1] > 1:
        d = sp_dist[:, 1:]
        w = np.exp(-d * 10)
        keep = w >= threshold
        src = np.broadcast_to(np.arange(sp_idx.shape[0])[:, None] + row0, d.shape)
        out.append((src[keep], sp_idx[:, 1:][keep].astype(np.int64), np.zeros(int(keep.sum()), np.int64), w[keep], d[keep], w[keep]))
    if mo_idx is not None and mo_idx.shape[1] > 1:
        s = mo_cos[:, 1:]
        keep = s >= np.float32(threshold)
        src = np.broadcast_to(np.arange(mo_idx.shape[0])[:, None] + row0, s.shape)
        sv = s[keep].astype(np.float64)
        out.append((src[keep], mo_idx[:, 1:][keep].astype(np.int64), np.ones(int(keep.sum()), np.int64), sv, sv, np.zeros_like(sv)))
    if not out:
        z = np.zeros(0)
        return dict(src=z.astype(np.int64), tgt=z.astype(np.int64), type=z.astype(np.int64), weight=z, f0=z, f1=z)
    cat = [np.concatenate([o[i] for o in out]) for i in range(6)]
    return dict(src=cat[0].astype(np.int64), tgt=cat[1], type=cat[2], weight=cat[3], f0=cat[4], f1=cat[5])


def candidate_edges_device(sp_idx, sp_dist, mo_idx, mo_cos, threshold, row0=0):
    """candidate_edges on device tensors (same emission order: spatial rows then morphological rows, row-major)."""
    dev = sp_idx.device if sp_idx is not None else mo_idx.device
    parts = []
    if sp_idx is not None and sp_idx.size(1) > 1:
        d = sp_dist[:, 1:]
        w = torch.exp(-d * 10)
        keep = w >= threshold
        src = (torch.arange(sp_idx.size(0), device=dev)[:, None] + row0).expand_as(d)
        parts.append((src[keep], sp_idx[:, 1:][keep].long(), torch.zeros(int(keep.sum()), dtype=torch.long, device=dev), w[keep], d[keep], w[keep]))
    if mo_idx is not None and mo_idx.size(1) > 1:
        sm = mo_cos[:, 1:]
        keep = sm >= torch.tensor(threshold, dtype=torch.float32, device=dev)       # float32 compare, like the reference's fp32 matrix
        src = (torch.arange(mo_idx.size(0), device=dev)[:, None] + row0).expand_as(sm)
        sv = sm[keep].double()
        parts.append((src[keep], mo_idx[:, 1:][keep].long(), torch.ones(int(keep.sum()), dtype=torch.long, device=dev), sv, sv, torch.zeros_like(sv)))
    names = ("src", "tgt", "type", "weight", "f0", "f1")
    if not parts:
        z = torch.zeros(0, device=dev)
        return {k: (z.long() if k in ("src", "tgt", "type") else z.double()) for k in names}
    return {k: torch.cat([p[i] for p in parts]) for i, k in enumerate(names)}


def remove_duplicate_edges_device(e, n):
    """remove_duplicate_edges on the device: two stable sorts give (key, -weight, first-seen) order; the winner of a key is the
    first row of its group, and groups are emitted in the order of their first occurrence."""
    m = e["src"].numel()
    if m == 0:
        return e
    a, b = torch.minimum(e["src"], e["tgt"]), torch.maximum(e["src"], e["tgt"])
    key = a * n + b
    by_w = torch.sort(-e["weight"], stable=True).indices
    by_k = torch.sort(key[by_w], stable=True).indices
    srt = by_w[by_k]
    ks = key[srt]
    first = torch.ones(m, dtype=torch.bool, device=key.device)
    first[1:] = ks[1:] != ks[:-1]
    winners = srt[first]
    pos_sorted = torch.sort(key, stable=True)
    kfirst = torch.ones(m, dtype=torch.bool, device=key.device)
    kfirst[1:] = pos_sorted.values[1:] != pos_sorted.values[:-1]
    first_pos = pos_sorted.indices[kfirst]                          # first occurrence of each key, keys ascending (like `winners`)
    emit = winners[torch.sort(first_pos, stable=True).indices]
    return {k: v[emit] for k, v in e.items()}


def remove_duplicate_edges(e, n):
    """_remove_duplicate_edges (:354-365): undirected key, strictly-greater weight replaces, first-seen position kept."""
    m = e["src"].shape[0]
    if m == 0:
        return e
    a, b = np.minimum(e["src"], e["tgt"]), np.maximum(e["src"], e["tgt"])
    key = a * np.int64(n) + b
    order = np.arange(m)
    # winner per key: max weight, earliest among equals; position: first occurrence of the key
    srt = np.lexsort((order, -e["weight"], key))
    ks = key[srt]
    first_of_group = np.ones(m, bool)
    first_of_group[1:] = ks[1:] != ks[:-1]
    winners = srt[first_of_group]
    _, first_pos = np.unique(key, return_index=True)            # sorted by key, like ks[first_of_group]
    emit = winners[np.argsort(first_pos, kind="stable")]
    return {k: v[emit] for k, v in e.items()}


class TissueGraphBuilder:
    """Same constructor as the reference (:50-59). The ImageNet/DINOv2 patch feature extractor of the reference
    (timm, network download, :85-137) is outside the hot path: patches must carry `features` (or fall back to the
    reference's 5 metadata placeholders)."""

    def __init__(self, feature_extractor="dinov2", spatial_k=8, morphological_k=16, edge_threshold=0.7,
                 hierarchical_levels=3, min_component_size=5, use_adaptive_threshold=True, device=None):
        self.feature_extractor = feature_extractor
        self.spatial_k, self.morphological_k, self.edge_threshold = spatial_k, morphological_k, edge_threshold
        self.hierarchical_levels, self.min_component_size = hierarchical_levels, min_component_size
        self.use_adaptive_threshold = use_adaptive_threshold
        self.logger = logging.getLogger(__name__)
        self.feature_dim = 768 if feature_extractor == "dinov2" else 512
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else None

    # ---------------------------------------------------------------------------------- nodes
    def _create_placeholder_features(self, patch):
        f = np.array([patch.tissue_percentage, patch.magnification / 40.0, float(patch.level),
                      np.log1p(patch.x / 1000.0), np.log1p(patch.y / 1000.0)])
        if len(f) < self.feature_dim:
            f = np.concatenate([f, np.zeros(self.feature_dim - len(f))])
        return f[:self.feature_dim]

    def _normalize_coordinates(self, x, y, metadata):
        if "dimensions" in metadata:
            w, h = metadata["dimensions"]
            return (x / w, y / h)
        return (x / 50000.0, y / 50000.0)

    def _create_nodes(self, slide_data, patch_images=None):
        if patch_images:
            raise NotImplementedError("patch image feature extraction (timm / CNN) is outside the B200 hot path; "
                                      "supply PatchInfo.features")
        nodes = []
        for p in slide_data.patches:
            f = p.features if p.features is not None else self._create_placeholder_features(p)
            nodes.append(GraphNode(p.patch_id, p, f, self._normalize_coordinates(p.x, p.y, slide_data.metadata)))
        return nodes

    # ---------------------------------------------------------------------------------- edges
    def edge_arrays(self, features, coords):
        """features float [n,d], coords float64 [n,2] (numpy). Returns the de-duplicated edge arrays in emission order."""
        if self.device is None:
            raise KernelError("TissueGraphBuilder needs a CUDA (B200) device; there is no CPU fallback")
        n = features.shape[0]
        f = torch.as_tensor(np.ascontiguousarray(features, dtype=np.float32)).to(self.device)
        c = torch.as_tensor(np.ascontiguousarray(coords, dtype=np.float64)).to(self.device)
        ks, km = min(self.spatial_k + 1, n), min(self.morphological_k + 1, n)
        sp_idx, sp_dist = knn2d(c, ks)
        mo_idx, mo_cos, _ = knn_cosine(f, km)
        # thresholds, undirected de-duplication and ordering stay on the device (index bookkeeping through torch's sort);
        # only the surviving edges come back to the host
        e = remove_duplicate_edges_device(candidate_edges_device(sp_idx, sp_dist, mo_idx, mo_cos, self.edge_threshold), n)
        return {k: v.cpu().numpy() for k, v in e.items()}

    # ---------------------------------------------------------------------------------- row-sharded construction (config 3)
    def shard_candidates(self, f, c, q0, nq):
        """Candidate edges of query rows [q0, q0+nq) — the rows ONE rank owns — against all n rows, in the reference's emission order.
        f float32 CUDA [n,d], c float64 CUDA [n,2]. Returns (spatial dict, morphological dict, near) where `near` counts the kNN
        pairs whose weight lies within 1e-6 of edge_threshold (the pairs whose keep/drop decision is sensitive to the last bits of
        the arithmetic; north_star asks for them to be counted and reported): [spatial, morphological]."""
        n = f.size(0)
        ks, km = min(self.spatial_k + 1, n), min(self.morphological_k + 1, n)
        sp_idx, sp_dist = knn2d(c, ks, q0=q0, nq=nq)
        mo_idx, mo_cos, _ = knn_cosine(f, km, q0=q0, nq=nq)
        thr = self.edge_threshold
        near_sp = int(((torch.exp(-sp_dist[:, 1:] * 10) - thr).abs() < 1e-6).sum().item()) if ks > 1 else 0
        near_mo = int(((mo_cos[:, 1:].double() - thr).abs() < 1e-6).sum().item()) if km > 1 else 0
        sp = candidate_edges_device(sp_idx, sp_dist, None, None, thr, row0=q0)
        mo = candidate_edges_device(None, None, mo_idx, mo_cos, thr, row0=q0)
        return sp, mo, (near_sp, near_mo)

    @staticmethod
    def merge_shards(shards, n):
        """shards: per rank (spatial dict, morphological dict) in rank order. The reference emits every spatial candidate (row-major)
        before any morphological one (:283-287), so the global list is [spatial of rank 0..P-1] + [morphological of rank 0..P-1];
        the undirected de-duplication (:354-365: strictly greater weight replaces, first-seen position kept) then runs on it."""
        names = ("src", "tgt", "type", "weight", "f0", "f1")
        parts = [sh[0] for sh in shards] + [sh[1] for sh in shards]
        cat = {k: torch.cat([p[k] for p in parts]) for k in names}
        return remove_duplicate_edges_device(cat, n)

    def edge_arrays_sharded(self, f, c, rank, world, group=None):
        """Edge construction with the query rows sharded over `world` ranks (one process per GPU, torch.distributed): every rank holds
        all n feature rows and coordinates, computes the 2-D and 768-d kNN of the rows it owns, and ONE exchange — an all-gather of
        the surviving candidate tuples — gives every rank the global candidate list, on which the de-duplication runs replicated.
        Returns (edge dict of device tensors in emission order, stats dict)."""
        import torch.distributed as dist
        n = f.size(0)
        rows = -(-n // world)
        q0, q1 = min(n, rank * rows), min(n, (rank + 1) * rows)
        sp, mo, near = self.shard_candidates(f, c, q0, q1 - q0)
        if world == 1:
            e = self.merge_shards([(sp, mo)], n)
            return e, {"near_threshold_spatial": near[0], "near_threshold_morphological": near[1], "candidates": int(sp["src"].numel() + mo["src"].numel())}
        dev = f.device
        cnt = torch.tensor([sp["src"].numel(), mo["src"].numel(), near[0], near[1]], dtype=torch.int64, device=dev)
        allc = torch.empty(world, 4, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, cnt, group=group)
        allc_h = allc.cpu()
        shards = [[None, None] for _ in range(world)]
        for which, part in enumerate((sp, mo)):
            mx = int(allc_h[:, which].max())
            ints = torch.zeros(max(mx, 1), 3, dtype=torch.int64, device=dev)
            flts = torch.zeros(max(mx, 1), 3, dtype=torch.float64, device=dev)
            m = part["src"].numel()
            if m:
                ints[:m] = torch.stack([part["src"], part["tgt"], part["type"]], dim=1)
                flts[:m] = torch.stack([part["weight"], part["f0"], part["f1"]], dim=1)
            gi = torch.empty(world * max(mx, 1), 3, dtype=torch.int64, device=dev)
            gf = torch.empty(world * max(mx, 1), 3, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(gi, ints, group=group)
            dist.all_gather_into_tensor(gf, flts, group=group)
            for r in range(world):
                k = int(allc_h[r, which])
                a = r * max(mx, 1)
                shards[r][which] = dict(src=gi[a:a + k, 0], tgt=gi[a:a + k, 1], type=gi[a:a + k, 2],
                                        weight=gf[a:a + k, 0], f0=gf[a:a + k, 1], f1=gf[a:a + k, 2])
        e = self.merge_shards([tuple(s_) for s_ in shards], n)
        return e, {"near_threshold_spatial": int(allc_h[:, 2].sum()), "near_threshold_morphological": int(allc_h[:, 3].sum()),
                   "candidates": int(allc_h[:, :2].sum())}

    def build_graph(self, slide_data, patch_images=None):
        nodes = self._create_nodes(slide_data, patch_images)
        if len(nodes) == 0:
            return self._create_empty_graph()
        ids = [nd.node_id for nd in nodes]
        if len(set(ids)) != len(ids):
            raise ValueError("patch ids must be unique")
        feats = np.stack([nd.features for nd in nodes])
        coords = np.array([nd.spatial_coords for nd in nodes], dtype=np.float64)
        e = self.edge_arrays(feats, coords)
        m = e["src"].shape[0]
        x = torch.tensor(feats, dtype=torch.float)
        pos = torch.tensor(coords, dtype=torch.float)
        if m > 0:
            ei = np.empty((2, 2 * m), dtype=np.int64)
            ei[0, 0::2], ei[1, 0::2] = e["src"], e["tgt"]                # (s,t) then (t,s), :400
            ei[0, 1::2], ei[1, 1::2] = e["tgt"], e["src"]
            attr = np.repeat(np.stack([e["f0"], e["f1"]], axis=1), 2, axis=0)
            et = np.repeat(e["type"], 2)
            edge_index, edge_attr, edge_type = torch.tensor(ei), torch.tensor(attr, dtype=torch.float), torch.tensor(et)
        else:
            edge_index = torch.empty((2, 0), dtype=torch.long)
            edge_attr = torch.empty((0, 1), dtype=torch.float)
            edge_type = torch.empty((0,), dtype=torch.long)
        data = Data(x=x, edge_index=edge_index, edge_attr=edge_attr, pos=pos, edge_type=edge_type, num_nodes=len(nodes))
        data.slide_id = slide_data.slide_id
        data.num_patches = len(slide_data.patches)
        data.slide_metadata = slide_data.metadata
        return data

    # ---------------------------------------------------------------------------------- hierarchy
    def create_hierarchical_graph(self, base_graph, levels=3):
        """preprocessing/tissue_graph_builder.py:449-471: [base, pooled, pooled of pooled, ...] with pool_ratio 0.5."""
        graphs, cur = [base_graph], base_graph
        for _ in range(1, levels):
            cur = self._pool_graph(cur, pool_ratio=0.5)
            graphs.append(cur)
        return graphs

    def _pool_graph(self, graph, pool_ratio=0.5):
        """:473-516. Keep the max(1, int(ratio*N)) nodes of highest degree (occurrences in either row of edge_index), ascending
        order, relabel and filter the edges (order kept). The reference computes the degrees with an O(N*E) Python loop and
        selects with torch.topk, whose choice among EQUAL degrees at the cut is implementation-defined; here the degrees are a
        device histogram and the exact radix select breaks ties towards the lower node index."""
        if self.device is None:
            raise KernelError("TissueGraphBuilder needs a CUDA (B200) device; there is no CPU fallback")
        from . import ops
        n = int(graph.num_nodes)
        k = max(1, int(n * pool_ratio))
        if k >= n:
            return graph
        dev = self.device
        ei = graph.edge_index.to(dev)
        if ei.numel() > 0:
            deg = torch.bincount(ei.reshape(-1), minlength=n).to(torch.float32)      # out-degree + in-degree
        else:
            deg = torch.zeros(n, dtype=torch.float32, device=dev)
        perm, node_map = ops.topk_select(deg.contiguous(), k)
        keep = perm.long()
        nm = node_map.long()
        host = graph.x.device
        if ei.numel() > 0:
            emask = (nm[ei[0]] >= 0) & (nm[ei[1]] >= 0)
            new_ei = nm[ei[:, emask]].to(host)
            ea = graph.edge_attr
            new_ea = ea[emask.to(ea.device)] if ea is not None and ea.numel() > 0 else torch.empty((0, 1))
        else:
            new_ei, new_ea = torch.empty((2, 0), dtype=torch.long), torch.empty((0, 1), dtype=torch.float)
        kh = keep.to(host)
        return Data(x=graph.x[kh], edge_index=new_ei, edge_attr=new_ea, pos=graph.pos[kh], num_nodes=k)

    def _create_empty_graph(self):
        return Data(x=torch.empty((0, self.feature_dim), dtype=torch.float), edge_index=torch.empty((2, 0), dtype=torch.long),
                    edge_attr=torch.empty((0, 1), dtype=torch.float), pos=torch.empty((0, 2), dtype=torch.float), num_nodes=0)
