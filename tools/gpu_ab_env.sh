#!/bin/bash
# A/B of an environment switch on ONE box, alternating:  gpu_ab_env.sh VAR "<bench args>"   (VAR=0 against VAR=1)
set -u
mkdir -p gpurun_out
VAR=$1; shift
for rep in 1 2; do for v in 0 1; do
  env $VAR=$v python bench.py "$@" > gpurun_out/abe_${v}_$rep.json 2> gpurun_out/abe_${v}_$rep.err
  python - $v $rep <<'PY'
import json, sys
v, rep = sys.argv[1:3]
d = json.loads(open(f"gpurun_out/abe_{v}_{rep}.json").read())
print(f"{v} rep {rep}: value {d['value']:.4f} {d['unit']}  ms_per_step {d['ms_per_step']:.1f}  clk {d['clocks']['sm_mhz'] if d.get('clocks') else None}")
PY
done; done
