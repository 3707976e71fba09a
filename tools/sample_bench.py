"""DiffusionLayer.sample (DDPM loop, 50 steps): eager kernel-by-kernel vs the CUDA-graph replay. Usage: python tools/sample_bench.py [n]"""
import os, sys, time, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgdm_histopath_lab_b200 import DiffusionLayer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
torch.manual_seed(0)
dl = DiffusionLayer(128, 256, num_timesteps=10).cuda().eval()
dev = torch.device("cuda")
for use_graph in (False, True):
    for _ in range(2):
        dl.sample((n, 128), dev, num_inference_steps=50, use_graph=use_graph)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        dl.sample((n, 128), dev, num_inference_steps=50, use_graph=use_graph)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"sample n={n} 50 steps use_graph={use_graph}: {dt*1e3:.2f} ms per call ({dt*1e6/50:.1f} us per step)")
