#!/bin/bash
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/gpu_latency_sweep.py C1 C2 C3 > gpurun_out/latency_sweep.jsonl 2> gpurun_out/latency_sweep.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/latency_sweep.jsonl; tail -3 gpurun_out/latency_sweep.err; cat gpurun_out/smoke.log
