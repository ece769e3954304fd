#!/bin/bash
# round-end checks on one GPU (gpurun): the whole -m gpu suite, smoke(), the N = 1 bench line.  See also gpu_run_scale.sh
# (1/2/4/8 GPUs + Hessian), gpu_ncu_round.sh (ncu evidence), gpu_latency_sweep.py, gpu_precision_study.py, gpu_overlap_probe.py.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cut -c1-200 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
