#!/bin/bash
# A/B of two builds on ONE box: pdb2reaction_b200/csrc/libumab_prev.so (copy of an earlier build) against the current
# libumab.so, alternating, same bench command; prints step time and the kernel families.  UMAB_LIB selects the build.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1; echo "parity rc=$?"
for rep in 1 2; do
UMAB_LIB=$PWD/pdb2reaction_b200/csrc/libumab_prev.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_prev_$rep.json 2> gpurun_out/ab_prev_$rep.err
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_new_$rep.json 2> gpurun_out/ab_new_$rep.err
done
python - <<'PY'
import json
for n in ("prev_1","new_1","prev_2","new_2"):
    d=json.loads(open(f"gpurun_out/ab_{n}.json").read()); f=d["kernel_families"]
    print(n, round(d["value"],2), "evals/s", round(d["ms_per_step"],1), "ms clk", d["clocks"]["sm_mhz"], "|", " ".join(f"{k}={v['ms_per_step']:.1f}" for k,v in f.items() if v['ms_per_step']>5))
PY
