#!/bin/bash
set -u
mkdir -p gpurun_out
( UMAB_DEBUG_GRAPH=1 python -m pytest tests/test_gpu_fastpath.py -m gpu -q -x ) > gpurun_out/pytest_fast.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fast.log
UMAB_DEBUG_GRAPH=1 python tools/gpu_latency_sweep.py C1 C2 C3 > gpurun_out/latency_sweep.jsonl 2> gpurun_out/latency_sweep.err
tail -30 gpurun_out/pytest_fast.log; cat gpurun_out/latency_sweep.jsonl; tail -5 gpurun_out/latency_sweep.err
