set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_calculator.py -m gpu -q -x -k "analytic or hessian" 2>&1 | tail -3
python bench.py --hessian --hessian-mode analytic --steps 1 > gpurun_out/hess_analytic_n1.json 2> gpurun_out/hess_analytic_n1.err; echo "rc=$?"
cut -c1-400 gpurun_out/hess_analytic_n1.json; tail -3 gpurun_out/hess_analytic_n1.err
