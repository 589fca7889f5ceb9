set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_image.py tests/test_gpu_step.py -q -k "prefetcher or training_driver or drop_in or e2e or main" > gpurun_out/pytest_gpu9.log 2>&1; tail -3 gpurun_out/pytest_gpu9.log
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline --no-loader-leg --no-reuse-leg > gpurun_out/bench_H_r01x.log 2>&1; tail -1 gpurun_out/bench_H_r01x.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
