#!/usr/bin/env python
"""bench.py — DGDM-Base forward nodes/sec on a synthetic 100k-node, 768-d, ~16-degree tile graph (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     the reference's CPU implementation of the path (oracle port of the
                                                           patched reference; the unpatched reference forward raises for every input)

One JSON line on rank 0. `value`: inputs resident in HBM, K whole forwards (CSR build included: the structure cache is dropped every step) timed with
CUDA events, max over ranks.
`e2e`: same metric through the public API with pinned HOST buffers (H2D of x/edge_index/pos and D2H of the graph
embedding inside the timed region, CSR rebuilt every step because the inputs are new tensors).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASE_CFG = dict(node_features=768, hidden_dims=[512, 256, 128], num_diffusion_steps=10, attention_heads=8, graph_layers=4)
LARGE_CFG = dict(node_features=1024, hidden_dims=[768, 512, 256, 128], num_diffusion_steps=20, attention_heads=16, graph_layers=6)
METRIC = "DGDM fwd nodes/sec (100k-node graph)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed `ncu --set full` summary under
    profiles/ (captured at this workload on one GPU); (None, None) when no capture names the kernel."""
    import glob
    import re
    best = (None, None)
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "attn_tc_*.txt"))):
        txt = open(path).read()
        if kernel not in txt:
            continue
        rd = re.search(r"dram__bytes_read\.sum\s+([0-9.]+)\s+(\w+)", txt)
        wr = re.search(r"dram__bytes_write\.sum\s+([0-9.]+)\s+(\w+)", txt)
        if rd and wr:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            best = (float(rd.group(1)) * mult.get(rd.group(2), 1.0) + float(wr.group(1)) * mult.get(wr.group(2), 1.0), os.path.relpath(path, ROOT))
    return best


def ncu_traffic_sum(kernel):
    """Sum of dram__bytes_read + dram__bytes_write over ALL launches of `kernel` in one forward, from the newest committed per-launch
    ncu table under profiles/ (profiles/spmm_dram_*.json, written by tools/ncu_dram_table.py); (None, None) when absent."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "spmm_dram_*.json")))
    if not paths:
        return None, None
    d = json.load(open(paths[-1]))
    rows = [r for r in d.get("launches", []) if kernel in r.get("kernel", "")]
    if not rows:
        return None, None
    return float(sum(r["dram_bytes"] for r in rows)), os.path.relpath(paths[-1], ROOT)


class ClockSampler:
    def __init__(self, dev):
        self.dev, self.rows, self._stop, self._t = dev, [], threading.Event(), None

    def start(self):
        def run():
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            while not self._stop.is_set():
                try:
                    o = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5).stdout.strip()
                    if o:
                        self.rows.append([c.strip() for c in o.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def _cpu_inputs(n_nodes):
    from dgdm_histopath_lab_b200 import DGDMModel, synth
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = DGDMModel(**BASE_CFG).eval()
    return synth.seeded_state_dict(m, 0), synth.tile_graph(n_nodes, 768, seed=1234)


def oracle_nodes_per_sec(n_nodes, reps, threads=None):
    """CPU baseline, kind "port": oracle/dgdm_oracle.py (restatement of the patched reference forward) on an n_nodes tile graph."""
    from oracle import dgdm_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd, d = _cpu_inputs(n_nodes)
    cfg = dict(hidden_dims=BASE_CFG["hidden_dims"], attention_heads=8, graph_layers=4)
    times = []
    with torch.no_grad():
        O.dgdm_forward(sd, cfg, d.x, d.edge_index, None, None, d.pos)
        for _ in range(reps):
            t0 = time.perf_counter()
            O.dgdm_forward(sd, cfg, d.x, d.edge_index, None, None, d.pos)
            times.append(time.perf_counter() - t0)
    return n_nodes / min(times), threads, times


def reference_nodes_per_sec(n_nodes, reps, threads=None):
    """CPU baseline, kind "reference": the reference's OWN files (staged unmodified under oracle/_ref by oracle/stage_ref.py, or the
    live tree in the build container) with the patch set P1-P7 applied at import time (the unpatched forward raises for every input,
    SURVEY.md 8-D), DGDMModel.forward through its public API on the same seeded weights and graph. Returns
    (nodes/s with monitor_operation stubbed, nodes/s as shipped, threads, times_stubbed) or None when no reference tree is available."""
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    ns = ref_loader.load(stub_monitor=False)
    sd, d = _cpu_inputs(n_nodes)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ns["DGDMModel"](**BASE_CFG).eval()
    m.load_state_dict(sd)
    rd = ns["Data"](x=d.x, edge_index=d.edge_index, pos=d.pos)
    res = {}
    import logging
    logging.disable(logging.CRITICAL)
    try:
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for stub in (True, False):
                ref_loader.set_monitor_stub(stub)
                m(rd, return_embeddings=True)
                ts = []
                for _ in range(reps if stub else max(1, min(reps, 2))):
                    t0 = time.perf_counter()
                    m(rd, return_embeddings=True)
                    ts.append(time.perf_counter() - t0)
                res[stub] = ts
    finally:
        logging.disable(logging.NOTSET)
    return n_nodes / min(res[True]), n_nodes / min(res[False]), threads, res[True]


def cpu_arm(n_nodes, reps):
    """The CPU arm of the comparison: the real reference when a tree is available (kind "reference"), else the oracle port."""
    r = reference_nodes_per_sec(n_nodes, reps)
    if r is not None:
        v, v_shipped, threads, times = r
        return {"value": v, "unit": "nodes/s", "cores": threads, "kind": "reference", "ms": 1e3 * min(times),
                "as_shipped_monitor_value": v_shipped,
                "sample": f"the reference's own DGDMModel.forward (files staged unmodified under oracle/_ref, patches P1-P7 applied at import, "
                          f"torch CPU, {threads} threads) on a {n_nodes}-node tile graph of the same generator, best of {len(times)} after 1 warm-up; "
                          f"`value` has the reference's monitor_operation (gc.collect per stage) stubbed, `as_shipped_monitor_value` keeps it; "
                          f"the 100k-node workload is infeasible on CPU (dense attention scores = 320 GB)"}, times
    v, threads, times = oracle_nodes_per_sec(n_nodes, reps)
    return {"value": v, "unit": "nodes/s", "cores": threads, "kind": "port", "ms": 1e3 * min(times),
            "sample": f"oracle port of the patched reference forward (P1-P7), {n_nodes}-node graph, best of {len(times)} after 1 warm-up "
                      f"(no reference tree staged under oracle/_ref)"}, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_nodes
    t0 = time.perf_counter()
    cpu, times = cpu_arm(n, max(1, args.steps))
    v = cpu["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "nodes/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * min(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"DGDM-Base forward, one {n}-node 768-d ~16-degree tile graph (bounded sample of the 100k-node workload: "
                                   f"the reference's dense N x N attention needs 320 GB at 100k nodes)", "nodes": n},
            "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def run_pretrain(args, rank, world, dev):
    """BASELINE.json configs[4]: self-supervised entity-masking pretrain step of DGDM-Base (models/dgdm_model.py:447-506:
    masking_ratio 0.15, cosine schedule, diffusion loss) — forward + backward + gradient all-reduce + AdamW step — slide-batch
    data-parallel: every rank trains on its own synthetic graph of --nodes nodes (global batch = one graph per GPU)."""
    import torch.distributed as dist
    from dgdm_histopath_lab_b200 import DGDMModel, synth
    from dgdm_histopath_lab_b200.train import allreduce_gradients
    n = args.nodes if args.nodes != 100000 else 50000
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = DGDMModel(**BASE_CFG)
    model.load_state_dict(synth.seeded_state_dict(model, 0))
    model = model.to(dev).train()
    data = synth.tile_graph(n, 768, seed=500 + rank).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, foreach=True)

    def step():
        opt.zero_grad(set_to_none=True)
        out = model.pretrain_step(data, mask_ratio=0.15)
        out["total_pretrain_loss"].backward()
        if world > 1:
            allreduce_gradients(model)
        opt.step()
        return out["total_pretrain_loss"].detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warm = max(1, min(args.steps, 5)), 2
    for _ in range(warm):
        loss = step()
    barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    l0 = ops_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops_launches() - l0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    gn = math.sqrt(sum(float(p.grad.double().pow(2).sum()) for p in model.parameters() if p.grad is not None))
    if rank == 0:
        per = ms / steps
        line = {"metric": "DGDM-Base pretrain step nodes/sec (fwd + bwd + grad all-reduce + AdamW)", "value": n * world / (per / 1e3), "unit": "nodes/s",
                "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"configs[4]: entity-masking pretrain step (mask_ratio 0.15, cosine schedule, 10 diffusion steps), DGDM-Base, one synthetic "
                                       f"{n}-node 768-d tile graph per rank, slide-batch data parallel over {world} GPU(s)", "nodes_per_rank": n,
                           "global_batch_graphs": world, "parallelism": "single GPU" if world == 1 else f"dp{world}: bucketed NCCL all-reduce of 7.3M fp32 gradients",
                           "backward": "hand-written kernels for dX GEMMs (tcgen05), transposed-CSR SpMM and the attention backward (fp32 SIMT recompute); "
                                       "dW via cuBLAS, row-wise ops via torch autograd"},
                "clocks": clocks, "gpu_launches": launches, "final_loss": float(loss), "grad_norm": gn,
                "reference_feasibility": f"the reference's dense attention needs {8.0 * n * n * 4 / 1e9:.0f} GB at {n} nodes (forward alone)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ops_launches():
    from dgdm_histopath_lab_b200 import ops
    return ops.LAUNCHES["kernels"]


def run_builder(args, rank, world, dev):
    """BASELINE.json configs[2]: TissueGraphBuilder edge construction (tissue_graph_builder.py:269-437: 8-NN on 2-D coordinates with
    w = exp(-10 d) >= 0.7, 16-NN on 768-d features kept when cosine >= 0.7, undirected de-duplication, (s,t),(t,s) emission) over
    --rows synthetic patch embeddings. One step = the whole construction: every rank all-gathers the feature rows it does not hold
    (NCCL), computes the kNN of the query rows it owns, one all-gather of the surviving candidate tuples, replicated de-duplication.
    Exactness: sampled query rows are re-ranked by the exact fp64 path (scikit-learn's arithmetic); the full edge list is compared
    with the scikit-learn oracle at n = 10k / 25k in tests/test_gpu_round2.py."""
    import torch.distributed as dist
    from dgdm_histopath_lab_b200 import TissueGraphBuilder, builder, synth
    n, d, nc = args.rows, 768, 4096
    g = torch.Generator(device=dev).manual_seed(1234)           # same seed on every rank -> identical rows (SURVEY 8(d) config-3 generator)
    centres = torch.nn.functional.normalize(torch.randn(nc, d, generator=g, device=dev), dim=1)
    sigma = 0.018 + 0.010 * torch.rand(nc, generator=g, device=dev)
    assign = torch.randint(0, nc, (n,), generator=g, device=dev)
    F = torch.empty(n, d, device=dev)
    for a in range(0, n, 65536):
        b = min(n, a + 65536)
        F[a:b] = centres[assign[a:b]] + sigma[assign[a:b], None] * torch.randn(b - a, d, generator=g, device=dev)
    xs, ys, dims = synth.jittered_grid_coords(n, seed=7)
    import numpy as np
    C = torch.as_tensor(np.stack([xs / dims[0], ys / dims[1]], axis=1)).to(dev)
    rows = -(-n // world)
    q0, q1 = min(n, rank * rows), min(n, (rank + 1) * rows)
    mine = torch.zeros(rows, d, device=dev)
    mine[: q1 - q0] = F[q0:q1]
    tb = TissueGraphBuilder(feature_extractor="none", device=dev)

    def step():
        if world > 1:                                           # each rank contributes the rows it holds; everybody needs all keys
            full = torch.empty(rows * world, d, device=dev)
            dist.all_gather_into_tensor(full, mine)
            f = full[:n]
        else:
            f = F
        return tb.edge_arrays_sharded(f, C, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warm = max(1, min(args.steps, 3)), 1
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    builder.KNN_STATS.update(rows=0, flagged=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        e, stats = step()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # exactness on sampled rows of this shard: the fp64 CUDA-core path (sklearn's ranking arithmetic), one row at a time
    check = 16
    idx, cos, _ = builder.knn_cosine(F, 17, q0=q0, nq=min(q1 - q0, 4096), mode="tc")
    bad = 0
    for r in torch.linspace(0, min(q1 - q0, 4096) - 1, check).long().tolist():
        i1, c1, _ = builder.knn_cosine(F, 17, q0=q0 + r, nq=1, mode="exact")
        bad += int(not (torch.equal(i1[0], idx[r]) and torch.equal(c1[0], cos[r])))
    st = torch.tensor([ms, wall * 1e3, float(bad), float(builder.KNN_STATS["flagged"])], device=dev, dtype=torch.float64)
    if world > 1:
        mx = st.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = st.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = st
    if rank == 0:
        per = mx[0].item() / steps
        m_und = int(e["src"].numel())
        types = torch.bincount(e["type"], minlength=2).tolist()
        line = {"metric": "TissueGraphBuilder edge construction rows/sec (768-d cosine + threshold 0.7)", "value": n / (per / 1e3), "unit": "rows/s",
                "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": per, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32 (ranking in f64, candidates on split-fp16 tcgen05)", "data": "synthetic",
                "config": {"workload": f"configs[2]: edge construction over {n} synthetic 768-d patch embeddings (4096 Gaussian clusters, within-cluster "
                                       f"cosine straddling 0.7) + jittered-grid coordinates, spatial_k=8, morphological_k=16, edge_threshold=0.7; query rows "
                                       f"sharded over {world} GPU(s)", "rows": n,
                           "parallelism": "single GPU" if world == 1 else f"query rows sharded over {world} ranks; NCCL all-gather of feature rows and of candidate tuples; replicated dedup"},
                "clocks": clocks, "wall_ms_per_step": mx[1].item() / steps,
                "edges_undirected": m_und, "edges_emitted": 2 * m_und, "edges_spatial": types[0], "edges_morphological": types[1],
                "near_threshold_pairs_within_1e-6": {"spatial": stats["near_threshold_spatial"], "morphological": stats["near_threshold_morphological"]},
                "candidate_tuples_exchanged": stats["candidates"],
                "gram_tflops_useful": 2.0 * n * n * d / (per / 1e3) / 1e12, "rows_flagged_and_recomputed_exactly": int(sm[3].item()),
                "sampled_rows_checked_vs_exact_fp64": check * world, "sampled_rows_mismatching": int(sm[2].item()),
                "reference_feasibility": f"the reference's dense cosine matrix alone is {4.0 * n * n / 1e12:.1f} TB at {n} rows"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nodes", type=int, default=100000)
    ap.add_argument("--ref-nodes", type=int, default=8000,
                    help="size of the CPU reference sample: the dense N x N attention of the reference needs 8*N^2*4 bytes (2 GB at 8k, 320 GB at 100k)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--rows", type=int, default=1000000, help="builder1m: number of synthetic 768-d patch embeddings")
    ap.add_argument("--workload", default="graph100k", choices=["graph100k", "batch4x50k", "large", "builder1m", "pretrain"],
                    help="graph100k = BASELINE.json's metric (one 100k-node tile graph); batch4x50k = configs[1] (4 x 50k-node graphs in one Batch); "
                         "large = configs[3]: DGDM-Large (configs/dgdm_large.yaml) on one --nodes graph (500000 in BASELINE.json), meant for torchrun; "
                         "pretrain = configs[4]: entity-masking pretrain step (fwd + bwd + grad all-reduce + AdamW) of DGDM-Base, one --nodes graph "
                         "(50000 in BASELINE.json) per rank; builder1m = configs[2]: TissueGraphBuilder edge construction over --rows 768-d embeddings, query rows sharded over the ranks")
    ap.add_argument("--cuda-graph", type=int, default=1, help="1: `value` replays the CUDA graph of the whole forward (CSR build included) captured by "
                                                              "DGDMModel.graphed(); 0: eager launches")
    ap.add_argument("--no-fused-halo", action="store_true", help="multi-GPU: NCCL all-gather per propagate instead of the GEMM-epilogue P2P/multicast stores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from dgdm_histopath_lab_b200 import DGDMModel, Data, _lib, ops, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()
    ops.device_check(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout must carry exactly ONE JSON line: NCCL prints its version banner to fd 1 when the first communicator is created,
        # so fd 1 points at stderr while the process group comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    if args.workload == "builder1m":
        return run_builder(args, rank, world, dev)
    if args.workload == "pretrain":
        return run_pretrain(args, rank, world, dev)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = DGDMModel(**(LARGE_CFG if args.workload == "large" else BASE_CFG)).eval()
    model.load_state_dict(synth.seeded_state_dict(model, 0))
    model = model.to(dev)
    feat = LARGE_CFG["node_features"] if args.workload == "large" else 768
    if args.workload == "batch4x50k":
        from dgdm_histopath_lab_b200 import Batch
        host = Batch.from_data_list([synth.tile_graph(args.nodes // 2, 768, seed=100 + i) for i in range(4)])
    elif args.workload == "large":
        host = synth.multimag_graph(args.nodes, feat, seed=1234)      # 3 magnification levels, parent / child links
    else:
        host = synth.tile_graph(args.nodes, feat, seed=1234)
    N, E = host.x.size(0), host.edge_index.size(1)
    edge_cut = synth.edge_cut_fraction(host.edge_index, N, world)
    host_batch = getattr(host, "batch", None)
    rows = None
    if world > 1:
        # 1-D node partition (parallel.py): every rank holds the graph structure and pos, and only its own rows of x
        from dgdm_histopath_lab_b200.parallel import even_bounds
        model.set_distributed(rank, world, max_nodes=N, fused=not args.no_fused_halo)
        b = even_bounds(N, world)
        rows = (b[rank], b[rank + 1])
        data = Data(x=host.x[rows[0]:rows[1]].to(dev), edge_index=host.edge_index.to(dev), pos=host.pos.to(dev), num_nodes=N)
        if host_batch is not None:
            data.batch = host_batch.to(dev)
    else:
        data = host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        # every step is a whole forward: the graph-structure cache (CSR by destination, GCN norm) is dropped so that the CSR build
        # is inside the timed region, as it is for a graph seen for the first time
        data.__dict__.pop("_dgdm_ctx", None)
        return model(data, return_embeddings=True)

    for _ in range(args.warmup):
        out = step_resident()
    barrier()
    graphed = None
    if args.cuda_graph and world == 1:
        try:
            graphed = model.graphed(data, return_embeddings=True, warmup=1, rebuild_structure=True)
            eager_step, step_resident = step_resident, graphed
            for _ in range(2):
                out = step_resident()
            barrier()
        except Exception as e:                       # capture unsupported in this configuration: keep the eager loop and say so
            print(f"[bench] CUDA-graph capture failed, timing eager launches: {e}", file=sys.stderr)
            graphed = None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.LAUNCHES["kernels"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.LAUNCHES["kernels"] - l0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = N * args.steps / (ms / 1e3)

    # ---- end to end through the public API with host buffers: every step copies that step's inputs from pinned host memory
    # (this rank's rows of x + the graph structure) and reads the graph embedding back; DevicePrefetcher stages step i+1 on
    # a side stream while step i computes, the CSR is rebuilt every step because the inputs are new tensors.
    e2e = None
    if not args.no_e2e:
        from dgdm_histopath_lab_b200.loader import DevicePrefetcher
        hg = Data(x=host.x.pin_memory(), edge_index=host.edge_index.pin_memory(), pos=host.pos.pin_memory())
        if host_batch is not None:
            hg.batch = host_batch.pin_memory()
        nx = N if rows is None else rows[1] - rows[0]
        h2d = nx * feat * 4 + hg.edge_index.numel() * 8 + hg.pos.numel() * 4

        # public serving loop: staging one graph ahead, a captured CUDA graph per staging slot at one GPU (eager launches when the model
        # is node-partitioned or --cuda-graph 0), validation and result read back on their own stream behind the next forward
        from dgdm_histopath_lab_b200.loader import GraphServer
        server = GraphServer(model, dev, use_graphs=bool(args.cuda_graph))

        def run_e2e(k):
            g = None
            for out in server.run([hg] * k, rows=rows):
                g = out["graph_embedding"].cpu()
            return g

        run_e2e(2)
        barrier()
        t0 = time.perf_counter()
        g = run_e2e(args.steps)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": N * args.steps / dt, "unit": "nodes/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": g.numel() * 4 + 32,                                     # result + the 32 validation bytes
               "ms_per_step": 1e3 * dt / args.steps, "staging": "pinned host -> device on a side stream, one step ahead",
               "api": "loader.GraphServer: " + ("CUDA-graph replay per staging slot" if server.replays else "eager launches (node-partitioned / --cuda-graph 0)")
                      + "; result + validation bytes read back on their own stream behind the next forward"}

    # ---- per-kernel-class timing pass (CUDA events after every C-ABI call on the launching stream; eager launches)
    pk = peaks()
    if graphed is not None:
        step_resident = eager_step
        launches = args.steps * graphed.launches
    step_resident()          # un-profiled eager pass first: the serving loop above captured its own graphs (private memory pools), so
    torch.cuda.synchronize() # the caching allocator may have to cudaMalloc fresh blocks on the first eager forward after it
    ops.profile_begin()
    step_resident()
    prof = ops.profile_end()
    meta = ops.take_profile_meta()
    att_ms = prof.get("dgdm_spatial_attention", [0, 0.0])[1]
    spmm_ms = prof.get("dgdm_spmm_gcn", [0, 0.0])[1]
    gemm_ms = prof.get("dgdm_gemm_f16x2", [0, 0.0])[1]
    nq = N if world == 1 else -(-N // world)
    att_flops = 4.0 * nq * N * 128 + 0.0                                      # QK^T + PV, per rank (query rows sharded)
    n_heads = 16 if args.workload == "large" else 8
    att_exp = float(n_heads) * nq * N
    spmm_bytes = sum((e + n) * c * 4 + n * c * 4 + (e + n) * 8 + (n + 1) * 4 for (n, e, c) in meta.get("spmm", []))
    gemm_flops = sum(2.0 * m * k * n_ for (m, k, n_) in meta.get("gemm", []))
    # Dominant kernel = the spatial attention core. It is bound by neither HBM nor the tensor pipe but by the XU (MUFU) pipe:
    # one exp2 per score plus the distance sqrt -> `bound: "mufu"` carries the fraction that applies (achieved / peak in G MUFU
    # ops/s at the SM clock sampled during the run: 148 SMs x 16 lanes/clk), and the tensor-pipe figure rides along in `tensor`.
    sm_clock_ghz = ((clocks or {}).get("sm_mhz") or 1965) / 1e3
    mufu_peak = 148 * 16 * sm_clock_ghz
    large = args.workload == "large"
    ops_per_score = ops.ATTN_MUFU_PER_SCORE                                   # exp2 + the share of the distance sqrt
    roof = {"kernel": "spatial_attn_tc2_kernel (+ prep / combine kernels of the same dgdm_spatial_attention call)"
                      + ("; DGDM-Large: 16 heads x 8 zero-padded to 16 dims, two 8-head passes" if large else ""),
            "bound": "mufu", "unit": "G MUFU ops/s", "ms": att_ms, "ops_per_score": ops_per_score,
            "achieved": ops_per_score * att_exp / (att_ms * 1e-3) / 1e9 if att_ms else None, "peak": mufu_peak,
            "peak_source": "148 SMs x 16 MUFU lanes/clk x sampled SM clock (not in MEASURED_PEAKS.json, which holds HBM and bf16 only)",
            "note": "exp2(MUFU)-bound: head_dim 16 gives K=16 / N=16 MMAs, the tensor pipe idles at ~20 %",
            "gexp_per_s": att_exp / (att_ms * 1e-3) / 1e9 if att_ms else None,
            "tensor": {"achieved": att_flops / (att_ms * 1e-3) / 1e12 if att_ms else None, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                       "peak_source": pk["src"]}}
    roof["frac"] = roof["achieved"] / roof["peak"] if roof["achieved"] else None
    if roof["tensor"]["achieved"]:
        roof["tensor"]["frac"] = roof["tensor"]["achieved"] / roof["tensor"]["peak"]
    roof["traffic"], roof["traffic_source"] = ncu_traffic("spatial_attn_tc2_kernel")
    roof_hbm = {"kernel": "spmm_gcn_kernel (all launches)", "bound": "hbm", "achieved": spmm_bytes / (spmm_ms * 1e-3) / 1e9 if spmm_ms else None,
                "peak": pk["hbm"], "unit": "GB/s", "ms": spmm_ms, "launches": len(meta.get("spmm", [])),
                "algorithmic_bytes": spmm_bytes}
    roof_hbm["traffic"], roof_hbm["traffic_source"] = ncu_traffic_sum("spmm_gcn_kernel")
    roof_hbm["frac"] = roof_hbm["achieved"] / roof_hbm["peak"] if roof_hbm["achieved"] else None
    roof_gemm = {"kernel": "gemm_f16x2_kernel (all launches)", "bound": "tensor", "achieved": 3 * gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                 "peak": pk["tf_sustained"], "unit": "TFLOP/s (3 MMA passes counted)", "ms": gemm_ms, "launches": len(meta.get("gemm", []))}
    roof_gemm["frac"] = roof_gemm["achieved"] / roof_gemm["peak"] if roof_gemm["achieved"] else None

    # ---- same-input anchor: the sizes the CPU reference can run (config 1 = 2 000 nodes, and --ref-nodes), GPU beside CPU on the
    # IDENTICAL graph and weights. GPU: eager forward (CSR rebuilt per call), and the CUDA-graph replay of the same forward.
    # NVLink volume of the halo exchange: every propagate all-gathers the projected [N_level, C] matrix; a rank receives the rows it
    # does not own. Only level 0 of the U-Net is partitioned (pooled levels are replicated after one [N,128] all-gather).
    halo_bytes = None
    if world > 1:
        nloc = rows[1] - rows[0]
        widths = [c for (n_, e_, c) in meta.get("spmm", []) if n_ == nloc]
        halo_bytes = int(sum((N - nloc) * c * 4 for c in widths) + (N - nloc) * (128 * 4 + 384 * 4 + 256 * 4))
    cpu, small = None, None
    if rank == 0 and world == 1 and args.workload == "graph100k" and not args.no_cpu_baseline:
        small = {}
        for n_small in sorted({2000, args.ref_nodes}):
            ds = synth.tile_graph(n_small, 768, seed=1234).to(dev)

            def timed(fn, reps=20):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / reps, 1e3 * (time.perf_counter() - t0) / reps

            def eager():
                ds.__dict__.pop("_dgdm_ctx", None)
                return model(ds, return_embeddings=True)
            ms_eager, wall_eager = timed(eager)
            run = model.graphed(ds, return_embeddings=True)
            ms_graph, wall_graph = timed(run)
            cpu_s, _ = cpu_arm(n_small, 3 if n_small <= 4000 else 2)
            small[str(n_small)] = {"nodes": n_small, "gpu_ms_eager": ms_eager, "gpu_ms_cuda_graph": ms_graph,
                                   "gpu_wall_ms_eager": wall_eager, "gpu_wall_ms_cuda_graph": wall_graph,
                                   "cpu_ms": cpu_s["ms"], "cpu_kind": cpu_s["kind"], "cpu_cores": cpu_s["cores"],
                                   "speedup_eager": cpu_s["ms"] / ms_eager, "speedup_cuda_graph": cpu_s["ms"] / ms_graph}
            if n_small == args.ref_nodes:
                cpu = cpu_s
    elif rank == 0 and not args.no_cpu_baseline:
        cpu, _ = cpu_arm(args.ref_nodes, 2)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": (f"DGDM-Base forward (eval, inference mode, return_embeddings) on one synthetic {N}-node 768-d tile graph, "
                                        f"E={E} directed edges, random-init seeded weights") if args.workload == "graph100k" else
                                       (f"DGDM-Large forward (1024-d nodes, hidden [768,512,256,128], 16 heads, 6 layers) on one synthetic {N}-node "
                                        f"multi-magnification graph (configs[3]: 5x/20x/40x nested grids 1:16:64, 8-neighbourhood + parent/child + 16 "
                                        f"morphological links), E={E} directed edges") if args.workload == "large" else
                                       (f"DGDM-Base forward on a Batch of 4 synthetic {N // 4}-node 768-d tile graphs (configs[1]), E={E} directed edges"),
                           "nodes": N, "edges": E,
                           "l2_policy": "inputs larger than L2 (x = 307 MB, activations > 200 MB per layer)",
                           "parallelism": "single GPU" if world == 1 else f"1-D node partition over {world} GPUs: features row-partitioned, NCCL all-gather "
                                          f"of projected rows per propagate (halo), context-parallel attention, replicated graph structure",
                           "edge_cut_fraction": edge_cut, "halo_bytes_received_per_rank_per_forward": halo_bytes,
                           "value_includes": "whole forward incl. CSR build (structure cache dropped every step)",
                           "cuda_graph": graphed is not None,
                           "attention_precise": ops.attention_precise_default(), "gemm": ops.gemm_mode(),
                           "halo": None if world == 1 else type(model._comm).__name__ + ("/multicast" if getattr(model._comm, "multicast", False) else "")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "roofline_hbm": roof_hbm, "roofline_gemm": roof_gemm,
                "cpu_baseline": cpu, "small_graph": small, "per_call_ms": {k: round(v[1], 4) for k, v in prof.items()},
                "graph_embedding_checksum": float(out["graph_embedding"].double().abs().sum().item())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
