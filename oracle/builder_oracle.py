"""TEST INFRASTRUCTURE — CPU restatement of TissueGraphBuilder's edge construction
(preprocessing/tissue_graph_builder.py:269-437) with patch P7 (edge features zero-padded to width 2).
Uses scikit-learn exactly as the reference does (NearestNeighbors, cosine_similarity — third-party arithmetic that
is not vendored in the reference; scikit-learn>=1.0 unpinned in requirements.txt:10, 1.9.0 in this image), and plain
Python loops for the edge logic, so it is only meant for small cases.

Pinned against the reference's own builder by tests/test_oracle_vs_reference.py and tests/golden/builder_*.npz.
"""
import numpy as np


def spatial_edges(coords, spatial_k, threshold):
    """:290-320. Returns list of (src, tgt, 'spatial', weight, [distance, weight])."""
    from sklearn.neighbors import NearestNeighbors
    n = coords.shape[0]
    nb = NearestNeighbors(n_neighbors=min(spatial_k + 1, n)).fit(coords)
    dist, idx = nb.kneighbors(coords)
    out = []
    for i in range(n):
        for j in range(1, idx.shape[1]):
            w = np.exp(-dist[i, j] * 10)
            if w >= threshold:
                out.append((i, int(idx[i, j]), 0, w, (dist[i, j], w)))
    return out


def morphological_edges(features, morph_k, threshold):
    """:322-352."""
    from sklearn.neighbors import NearestNeighbors
    from sklearn.metrics.pairwise import cosine_similarity
    n = features.shape[0]
    sim = cosine_similarity(features)
    nb = NearestNeighbors(n_neighbors=min(morph_k + 1, n)).fit(features)
    _, idx = nb.kneighbors(features)
    out = []
    for i in range(n):
        for j in range(1, idx.shape[1]):
            s = sim[i, idx[i, j]]
            if s >= threshold:
                out.append((i, int(idx[i, j]), 1, s, (s, 0.0)))
    return out


def dedup(edges):
    """:354-365 — undirected key, strictly greater weight replaces, dict insertion order kept."""
    d = {}
    for e in edges:
        key = (min(e[0], e[1]), max(e[0], e[1]))
        if key not in d or e[3] > d[key][3]:
            d[key] = e
    return list(d.values())


def build_edges(features, coords, spatial_k=8, morph_k=16, threshold=0.7):
    """Returns (edge_index int64 [2,2M], edge_attr float32 [2M,2], edge_type int64 [2M]) as _to_pytorch_geometric emits."""
    edges = dedup(spatial_edges(coords, spatial_k, threshold) + morphological_edges(features, morph_k, threshold))
    ei, ea, et = [], [], []
    for (s, t, ty, w, f) in edges:
        ei += [[s, t], [t, s]]
        ea += [f, f]
        et += [ty, ty]
    if not ei:
        return np.zeros((2, 0), np.int64), np.zeros((0, 2), np.float32), np.zeros((0,), np.int64)
    return np.array(ei, dtype=np.int64).T, np.array(ea, dtype=np.float32), np.array(et, dtype=np.int64)
