#!/bin/bash
# ncu evidence of one round, run on the GPU box (gpurun): launch list of one step + --set full captures of the GEMM and of
# the memory-bound edge kernels, on a 6-image C4 sub-batch (one closed chunk of ~620k edges).  Outputs in gpurun_out/.
set -u
OUT=gpurun_out
CMD="python bench.py --images 6 --steps 1 --warmup 0 --no-cpu-baseline"
$CMD > $OUT/ncu_plain.json 2> $OUT/ncu_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_final.csv $CMD > $OUT/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_pair -s 60 -c 27 -f -o $OUT/prof_pair $CMD > $OUT/ncu_p.log 2>&1
echo "pair capture rc=$?"
ncu -i $OUT/prof_pair.ncu-rep --page raw --csv > $OUT/prof_pair_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on \
    -k regex:"gather_rotate|rotate_back|combine_gate|ln_silu" -s 40 -c 40 -f -o $OUT/prof_edge3 $CMD > $OUT/ncu_e.log 2>&1
echo "edge capture rc=$?"
ncu -i $OUT/prof_edge3.ncu-rep --page raw --csv > $OUT/prof_edge3_raw.csv 2>/dev/null
ls -la $OUT/prof_pair* $OUT/prof_edge3* $OUT/launches_final.csv
rm -f $OUT/prof_edge3.ncu-rep        # keep the (smaller) GEMM report for the source page, the raw CSVs for everything else
