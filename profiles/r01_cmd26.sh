set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu4.log 2>&1; tail -4 gpurun_out/pytest_gpu4.log
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H_pow2.md > gpurun_out/bench_H_pow2.log 2>&1; tail -1 gpurun_out/bench_H_pow2.log | cut -c1-260
SIVAE_LIB_PATH=$PWD/profiles/_ab_prev.so timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H_prev.md > gpurun_out/bench_H_prev.log 2>&1; tail -1 gpurun_out/bench_H_prev.log | cut -c1-260
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H_pow2b.log 2>&1; tail -1 gpurun_out/bench_H_pow2b.log | cut -c1-260
