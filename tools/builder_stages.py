import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from dgdm_histopath_lab_b200 import builder, synth
from dgdm_histopath_lab_b200.builder import knn2d, knn_cosine, candidate_edges, remove_duplicate_edges
n = 100000
feats = synth.clustered_features(n, d=768, num_clusters=max(16, n // 244), seed=3)
xs, ys, dims = synth.jittered_grid_coords(n, seed=3, extent=100000)
coords = np.stack([xs / dims[0], ys / dims[1]], 1).astype(np.float64)
dev = "cuda"
def T(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); print(f"{name:28s} {1e3*(time.perf_counter()-t0):9.1f} ms"); return r
f = T("H2D features", lambda: torch.as_tensor(np.ascontiguousarray(feats, dtype=np.float32)).to(dev))
c = T("H2D coords", lambda: torch.as_tensor(coords).to(dev))
knn2d(c[:4096].contiguous(), 9); knn_cosine(f[:8192].contiguous(), 17)
sp = T("knn2d (brute fp64)", lambda: knn2d(c, 9))
mo = T("knn_cosine (tc)", lambda: knn_cosine(f, 17))
arrs = T("D2H results", lambda: (sp[0].cpu().numpy(), sp[1].cpu().numpy(), mo[0].cpu().numpy(), mo[1].cpu().numpy()))
e = T("candidate_edges (numpy)", lambda: candidate_edges(*arrs, 0.7))
print("candidates", e["src"].shape[0])
r = T("remove_duplicate_edges", lambda: remove_duplicate_edges(e, n))
from dgdm_histopath_lab_b200.builder import candidate_edges_device, remove_duplicate_edges_device
ed = T("candidate_edges (device)", lambda: candidate_edges_device(sp[0], sp[1], mo[0], mo[1], 0.7))
rd = T("remove_duplicates (device)", lambda: remove_duplicate_edges_device(ed, n))
print("kept (device)", rd["src"].numel())
print("kept", r["src"].shape[0])
