#!/bin/bash
set -u
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_parity.py tests/test_gpu_calculator.py tests/test_gpu_large.py tests/test_gpu_fastpath.py tests/test_gpu_refwrap.py -m gpu -q -x ) > gpurun_out/pytest_torque.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_torque.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_torque.json 2> gpurun_out/bench_torque.err
tail -12 gpurun_out/pytest_torque.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_torque.json").read())
f=d["kernel_families"]
print(round(d["value"],2), "evals/s", round(d["ms_per_step"],1), "ms; clk", d["clocks"]["sm_mhz"])
for k in ("gemm","gather_rotate_bwd","rotate_back_bwd","gather_rotate_scale","combine_gate_bwd","combine_gate_fwd","rotate_back_reduce","ln_silu"):
    print(f"  {k:22s} {f[k]['ms_per_step']:7.2f} ms  {f[k]['rate']:8.1f} {f[k]['rate_unit']}")
PY
