#!/bin/bash
# A/B/C of several builds on ONE box, alternating: usage  gpu_ab_libs.sh name1=path1 name2=path2 ...   (UMAB_LIB selects the build)
set -u
mkdir -p gpurun_out
for rep in 1 2; do
for kv in "$@"; do
  n=${kv%%=*}; p=${kv#*=}
  UMAB_LIB=$PWD/$p python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/abl_${n}_$rep.json 2> gpurun_out/abl_${n}_$rep.err
done
done
python - "$@" <<'PY'
import json, sys
for rep in (1, 2):
    for kv in sys.argv[1:]:
        n = kv.split("=")[0]
        try:
            d = json.loads(open(f"gpurun_out/abl_{n}_{rep}.json").read()); f = d["kernel_families"]
            print(n, rep, round(d["ms_per_step"], 1), "ms clk", d["clocks"]["sm_mhz"], "|", " ".join(f"{k}={v['ms_per_step']:.1f}" for k, v in f.items() if v['ms_per_step'] > 5))
        except Exception as e:
            print(n, rep, "FAILED", e)
PY
