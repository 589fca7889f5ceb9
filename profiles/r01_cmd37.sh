set -x
mkdir -p gpurun_out
timeout 300 python bench.py --config H --steps 20 --warmup 3 --no-cpu-baseline --no-loader-leg --no-reuse-leg --e2e-sweep > gpurun_out/bench_H_e2e_sweep.log 2>&1; tail -1 gpurun_out/bench_H_e2e_sweep.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
