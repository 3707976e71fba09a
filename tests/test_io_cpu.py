"""CPU: graph file formats of SURVEY 8(f) rank 2 (reference data/dataset.py:190-240): .dgdm round trip incl. the stored CSR,
.pt / .pkl loading, host CSR == the structure the oracle's message passing implies."""
import os
import pickle

import numpy as np
import torch

import dgdm_histopath_lab_b200 as pkg
from dgdm_histopath_lab_b200 import io as gio, synth


def _graph():
    d = synth.tile_graph(301, 24, seed=3)
    d.edge_attr = torch.rand(d.edge_index.size(1), 2, generator=torch.Generator().manual_seed(1))
    d.y = torch.tensor([2])
    return d


def test_dgdm_round_trip_and_alignment(tmp_path):
    d = _graph()
    f = pkg.save_graph(d, str(tmp_path / "g.dgdm"))
    g = pkg.load_graph(f)
    for k in ("x", "edge_index", "pos", "edge_attr", "y"):
        assert torch.equal(getattr(g, k), getattr(d, k)), k
    rowptr, col, eid = g.csr_host
    ei = d.edge_index.numpy()
    assert rowptr[0] == 0 and rowptr[-1] == ei.shape[1]
    for v in (0, 17, 300):                       # in-edges of v in ORIGINAL edge order
        want = np.nonzero(ei[1] == v)[0]
        a, b = int(rowptr[v]), int(rowptr[v + 1])
        assert np.array_equal(eid[a:b].numpy(), want) and np.array_equal(col[a:b].numpy(), ei[0][want])
    g2 = pkg.load_graph(pkg.save_graph(d, str(tmp_path / "nocsr.dgdm"), with_csr=False))
    assert not hasattr(g2, "csr_host") and torch.equal(g2.x, d.x)


def test_reference_formats_pt_and_pkl(tmp_path):
    d = _graph()
    torch.save(d, tmp_path / "g.pt")
    with open(tmp_path / "g.pkl", "wb") as fh:
        pickle.dump(d, fh)
    for name in ("g.pt", "g.pkl"):
        g = pkg.load_graph(str(tmp_path / name))
        assert torch.equal(g.x, d.x) and torch.equal(g.edge_index, d.edge_index)
    ds = pkg.GraphFileDataset(str(tmp_path))
    assert len(ds) == 2 and torch.equal(ds[0].x, d.x)
    try:
        pkg.load_graph(str(tmp_path / "g.txt"))
        assert False
    except ValueError as e:
        assert "Unsupported graph format" in str(e)
