#!/bin/bash
# Runs the GPU parity suites file by file (separate processes, bounded by timeout) so that one faulting kernel cannot hide
# the results of the others. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in test_gpu_kernels test_gpu_model; do
  timeout -k 10 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 30 gpurun_out/test_gpu_kernels.log
tail -n 30 gpurun_out/test_gpu_model.log
tail -n 5 gpurun_out/smoke.log
