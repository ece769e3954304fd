#!/bin/bash
# A/B on ONE box: the build before the torque adjoint (libumab_prev.so, commit 6a383d4) against the current build
set -u
mkdir -p gpurun_out
for rep in 1 2; do
# build it first: git archive <commit> | tar -x -C /tmp/prev; (cd /tmp/prev && python -m pdb2reaction_b200.csrc.build); cp .../libumab.so here as libumab_prev.so
UMAB_LIB=$PWD/pdb2reaction_b200/csrc/libumab_prev.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_prev_$rep.json 2> gpurun_out/ab_prev_$rep.err
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_new_$rep.json 2> gpurun_out/ab_new_$rep.err
done
python - <<'PY'
import json
for n in ("prev_1","new_1","prev_2","new_2"):
    d=json.loads(open(f"gpurun_out/ab_{n}.json").read()); f=d["kernel_families"]
    print(n, round(d["value"],2), "evals/s", round(d["ms_per_step"],1), "ms clk", d["clocks"]["sm_mhz"], "| gather_bwd", round(f["gather_rotate_bwd"]["ms_per_step"],1), "rotback_bwd", round(f["rotate_back_bwd"]["ms_per_step"],1), "gemm", round(f["gemm"]["ms_per_step"],1))
PY
