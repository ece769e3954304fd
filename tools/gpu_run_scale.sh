#!/bin/bash
# strong-scaled C4 string + Hessian on 1/2/4/8 GPUs of one box
set -u
mkdir -p gpurun_out
P=29500
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    python bench.py --gpus 1 --hessian --hessian-mode fd --steps 1 > gpurun_out/hessian_n$N.json 2> gpurun_out/hessian_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+10+N)) bench.py --gpus $N --hessian --hessian-mode fd --steps 1 > gpurun_out/hessian_n$N.json 2> gpurun_out/hessian_n$N.err
  fi
  echo "N=$N: $(cut -c1-160 gpurun_out/scale_n$N.json)"; echo "   H: $(cut -c1-120 gpurun_out/hessian_n$N.json)"
done
