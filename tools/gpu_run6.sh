#!/bin/bash
set -u
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_parity.py tests/test_gpu_gemm.py -m gpu -q -x ) > gpurun_out/pytest_fuse.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fuse.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
UMAB_FUSE_GATE=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err
tail -6 gpurun_out/pytest_fuse.log
python - <<'PY'
import json
for n in ("fused","unfused"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.json").read())
        f=d["kernel_families"]
        print(n, round(d["value"],2), "evals/s", round(d["ms_per_step"],1), "ms; gemm", round(f["gemm"]["ms_per_step"],1), "combine_fwd", round(f["combine_gate_fwd"]["ms_per_step"],1), "clk", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed", e)
PY
