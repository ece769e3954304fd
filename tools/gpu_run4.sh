#!/bin/bash
set -u
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_fastpath.py -m gpu -q -x ) > gpurun_out/pytest_fast.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fast.log
for c in C1 C3; do
UMAB_CUDA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_single_$c.csv python tools/gpu_single_call.py $c 3 > gpurun_out/ncu_single_$c.log 2>&1
done
tail -8 gpurun_out/pytest_fast.log; wc -l gpurun_out/launches_single_*.csv
