"""Dev tool: cycle accounting of the v2 attention kernel (needs a library built with -DT2_PROF=1, see tools/attn_ab.sh).
Prints where one softmax warp and one MMA-issuing thread of CTA (5, 1) spend their cycles."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import _lib, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
if len(sys.argv) > 2 and sys.argv[2] == "model":
    # the attention call of a real DGDM-Base forward (bench.py's model and graph): q / k statistics of the model, not randn
    import warnings
    from dgdm_histopath_lab_b200 import DGDMModel, synth
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = DGDMModel(node_features=768, hidden_dims=[512, 256, 128], num_diffusion_steps=10, attention_heads=8, graph_layers=4).eval()
    model.load_state_dict(synth.seeded_state_dict(model, 0))
    model = model.cuda()
    data = synth.tile_graph(n, 768, seed=1234).to("cuda")
    for _ in range(2):
        model(data, mode="inference", return_embeddings=True)
else:
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(n, 384, generator=g).cuda()
    pos = torch.rand(n, 2, generator=g).cuda()
    seg = torch.tensor([0, n], dtype=torch.int32, device="cuda")
    for _ in range(3):
        o = ops.spatial_attention(qkv, pos, seg, 1, 8)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * 32)()
lib.dgdm_debug_attn_prof.restype = C.c_int
assert lib.dgdm_debug_attn_prof(buf) == 0
v = list(buf)
nblk = max(v[9], 1)
units = 4 * nblk
names = ["wait full (per block)", "distance + arrive empty (per block)", "wait::ld at unit top", "need check + chunk a + st", "wait S(u+1) + ld issue",
         "chunk b + st issue", "wait::st + arrive + ld issue"]
print(f"softmax warp 0 of CTA (5,1): {nblk} blocks, total {v[8]} cycles = {v[8] / units:.0f} per unit")
for i, nm in enumerate(names):
    print(f"  {nm:40s} {v[i] / units:8.1f} cycles per unit")
print(f"  row-maximum pass taken in {v[7]} of {units} units")
names = ["first two QK issues", "wait P(u)", "PV + l issue", "QK(u+2) issue (incl. wait full)"]
print(f"MMA issuer g=0: total {v[24]} cycles = {v[24] / units:.0f} per unit")
for i, nm in enumerate(names):
    print(f"  {nm:40s} {v[16 + i] / units:8.1f} cycles per unit")
