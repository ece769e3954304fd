#!/bin/bash
# ncu --set full capture of the memory-bound edge kernels only (see tools/gpu_ncu_round.sh for the whole set)
set -u
OUT=gpurun_out
CMD="python bench.py --images 6 --steps 1 --warmup 0 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on \
    -k regex:"gather_rotate|rotate_back|combine_gate|ln_silu" -s 40 -c 40 -f -o $OUT/prof_edge4 $CMD > $OUT/ncu_e4.log 2>&1
echo "edge capture rc=$?"
ncu -i $OUT/prof_edge4.ncu-rep --page raw --csv > $OUT/prof_edge4_raw.csv 2>/dev/null
rm -f $OUT/prof_edge4.ncu-rep
ls -la $OUT/prof_edge4_raw.csv
