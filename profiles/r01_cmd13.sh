set -x
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_all.log 2>&1; tail -4 gpurun_out/pytest_all.log
SIVAE_TC_2CTA=2 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_2cta.log 2>&1; tail -2 gpurun_out/pytest_2cta.log
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H20.md > gpurun_out/bench_H20.log 2>&1; tail -1 gpurun_out/bench_H20.log | cut -c1-200
