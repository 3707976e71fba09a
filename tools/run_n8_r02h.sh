#!/bin/bash
# Multi-GPU evidence for the r02h kernels: headline scaling point (+ DGDM-Large config 4 at 8 GPUs). Usage: bash tools/run_n8_r02h.sh TAG NGPU
TAG=${1:-r02h}; N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "graph100k exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_n$N.json")); print(d["n_gpus"], round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["ms_per_step"],3), d["config"]["halo"], {k:round(v,2) for k,v in d["per_call_ms"].items()})
PY
if [ "$N" == "8" ]; then
timeout 300 $TR --master-port 29523 bench.py --gpus $N --workload large --nodes 500000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_large_${TAG}_n$N.json 2> gpurun_out/bench_large_${TAG}_n$N.err; echo "large exit $?"; tail -3 gpurun_out/bench_large_${TAG}_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_large_${TAG}_n$N.json")); print(d["n_gpus"], round(d["ms_per_step"],3), round(d["value"]), d["config"]["edge_cut_fraction"], {k:round(v,2) for k,v in d["per_call_ms"].items()})
PY
fi
