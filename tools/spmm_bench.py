"""SpMM (GCN propagate) timing at the BASELINE graph (100k nodes, 1.6M edges) for the three encoder widths; rotating inputs > L2.
    DGDM_SPMM_UNROLL=2|4|8 python tools/spmm_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import ops, synth  # noqa: E402

d = synth.tile_graph(100000, 8, seed=1234).to("cuda")
csr = ops.csr_build(d.edge_index, 100000)
E = d.edge_index.size(1)
res = []
for C in (512, 256, 128):
    hs = [torch.randn(100000, C, device="cuda") for _ in range(3)]
    bias = torch.randn(C, device="cuda")
    for h in hs:
        ops.spmm_gcn(csr, h, bias=bias, act="gelu")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    a.record()
    for i in range(reps):
        ops.spmm_gcn(csr, hs[i % 3], bias=bias, act="gelu")
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    byt = (E + 100000) * C * 4 + 100000 * C * 4 + (E + 100000) * 8
    res.append(f"C={C}: {us:7.1f} us  {byt / us / 1e3:7.0f} GB/s algorithmic")
print(f"unroll={os.environ.get('DGDM_SPMM_UNROLL', 'default')}  " + "   ".join(res))
