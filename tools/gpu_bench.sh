#!/bin/bash
# bench + ncu evidence for one GPU. Usage: bash tools/gpu_bench.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench_$TAG.json; tail -n 5 gpurun_out/bench_$TAG.err
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout -k 10 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches exit $?"
timeout -k 10 600 $CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:spatial_attn -s 3 -c 1 -f -o gpurun_out/attn_$TAG $CMD > gpurun_out/ncu_attn_$TAG.log 2>&1
echo "ncu attn exit $?"
ls -la gpurun_out | tail -20
