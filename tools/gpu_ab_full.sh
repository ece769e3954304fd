#!/bin/bash
# the whole -m gpu suite, then an A/B of two builds on ONE box, alternating: pdb2reaction_b200/csrc/libumab_prev.so (a copy
# of an earlier build, placed there by hand) against the current libumab.so; UMAB_LIB selects the build
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
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
