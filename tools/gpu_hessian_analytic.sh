#!/bin/bash
# analytic-Hessian checks on one GPU: the calculator tests that touch the dual-number path, then the C3 analytic line
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_calculator.py tests/test_gpu_fastpath.py -m gpu -q -x 2>&1 | tail -15
python bench.py --hessian --hessian-mode analytic --steps 1 > gpurun_out/hess_analytic_n1.json 2> gpurun_out/hess_analytic_n1.err; echo "rc=$?"
cut -c1-120 gpurun_out/hess_analytic_n1.json; tail -3 gpurun_out/hess_analytic_n1.err
