#!/bin/bash
# round 2, GPU run 1: tests, smoke, bench (N=1), precision study, compute-sanitizer, config sweep
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
python tools/gpu_precision_study.py > gpurun_out/precision_study.log 2>&1; echo "rc=$?" >> gpurun_out/precision_study.log
python tools/gpu_config_sweep.py C1 C2 C3 C5 > gpurun_out/config_sweep.jsonl 2> gpurun_out/config_sweep.err
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize_case.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_sanitize_case.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/gpu_sanitize_case.py > gpurun_out/sanitizer_synccheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_synccheck.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_short.json 2> gpurun_out/bench_ref_short.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_n1.json | cut -c1-600
