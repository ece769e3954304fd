#!/bin/bash
# ncu evidence of one round, run on the GPU box (gpurun): launch list of one step + --set full captures of the GEMM, of
# the memory-bound edge kernels (on a 6-image C4 sub-batch: one closed chunk of ~620k edges) and of the split-K cluster
# GEMM (20-atom call).  Outputs in gpurun_out/.  Profiling switches the engine to its synchronising path (prof.on).
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --images 6 --steps 1 --warmup 0 --no-cpu-baseline"
$CMD > $OUT/ncu_plain.json 2> $OUT/ncu_plain.err || { echo "plain run failed"; tail -5 $OUT/ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_final.csv $CMD > $OUT/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_pair -s 60 -c 27 -f -o $OUT/prof_pair $CMD > $OUT/ncu_p.log 2>&1
echo "pair capture rc=$?"
ncu -i $OUT/prof_pair.ncu-rep --page raw --csv > $OUT/prof_pair_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on \
    -k regex:"gather_rotate|rotate_back|combine_gate|ln_silu" -s 40 -c 40 -f -o $OUT/prof_edge $CMD > $OUT/ncu_e.log 2>&1
echo "edge capture rc=$?"
ncu -i $OUT/prof_edge.ncu-rep --page raw --csv > $OUT/prof_edge_raw.csv 2>/dev/null
UMAB_CUDA_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:gemm_simt_splitk -s 120 -c 20 -f -o $OUT/prof_splitk \
    python tools/gpu_single_call.py C1 3 > $OUT/ncu_s.log 2>&1
echo "split-K capture rc=$?"
ncu -i $OUT/prof_splitk.ncu-rep --page raw --csv > $OUT/prof_splitk_raw.csv 2>/dev/null
UMAB_CUDA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_single_C1_splitk.csv python tools/gpu_single_call.py C1 3 > $OUT/ncu_s1.log 2>&1
ls -la $OUT/prof_pair* $OUT/prof_edge* $OUT/prof_splitk* $OUT/launches_final.csv
rm -f $OUT/prof_edge.ncu-rep $OUT/prof_splitk.ncu-rep   # keep the (smaller) GEMM report for the source page, the raw CSVs for everything else
