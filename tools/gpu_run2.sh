#!/bin/bash
# round 2, GPU run 2: full GPU test-suite (no -x), latency sweep, bench N=1
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/gpu_latency_sweep.py C1 C2 C3 > gpurun_out/latency_sweep.jsonl 2> gpurun_out/latency_sweep.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/latency_sweep.jsonl; tail -3 gpurun_out/latency_sweep.err; cut -c1-300 gpurun_out/bench_n1.json
