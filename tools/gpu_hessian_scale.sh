#!/bin/bash
# analytic (default) or fd Hessian of C3 on the GPU counts given:  gpu_hessian_scale.sh analytic 1 2   (one box with max(N) GPUs)
set -u
mkdir -p gpurun_out
MODE=${1:-analytic}; shift
P=29600
for N in "$@"; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --hessian --hessian-mode $MODE --steps 1 > gpurun_out/hessian_${MODE}_n$N.json 2> gpurun_out/hessian_${MODE}_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --hessian --hessian-mode $MODE --steps 1 > gpurun_out/hessian_${MODE}_n$N.json 2> gpurun_out/hessian_${MODE}_n$N.err
  fi
  echo "N=$N rc=$?: $(cut -c1-110 gpurun_out/hessian_${MODE}_n$N.json)"; tail -2 gpurun_out/hessian_${MODE}_n$N.err
done
