#!/bin/bash
# end-of-round evidence on one GPU: the round checks, then the ncu launch list of one 6-image step of the final code
set -u
bash tools/gpu_round_checks.sh
CMD="python bench.py --images 6 --steps 1 --warmup 0 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_final.csv
