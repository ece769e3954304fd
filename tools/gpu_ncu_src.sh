#!/bin/bash
# ncu --set full WITH the source page of the three adjoint / reduce edge kernels that sit furthest below the HBM roofline
# (the .ncu-rep comes back: read it with  ncu -i rep --page source --csv).  6-image C4 sub-batch, one closed chunk.
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --images 6 --steps 1 --warmup 0 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on \
    -k regex:"rotate_back_bwd|gather_rotate_bwd_half|rotate_back_reduce" -s 3 -c 6 -f -o $OUT/prof_src $CMD > $OUT/ncu_src.log 2>&1
echo "capture rc=$?"
ncu -i $OUT/prof_src.ncu-rep --page raw --csv > $OUT/prof_src_raw.csv 2>/dev/null
ls -la $OUT/prof_src*
