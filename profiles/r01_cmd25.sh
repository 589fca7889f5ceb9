set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu3.log 2>&1; tail -4 gpurun_out/pytest_gpu3.log
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H_mask.md > gpurun_out/bench_H_mask.log 2>&1; tail -1 gpurun_out/bench_H_mask.log | cut -c1-260
SIVAE_BN_MASK=0 timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H_nomask.md > gpurun_out/bench_H_nomask.log 2>&1; tail -1 gpurun_out/bench_H_nomask.log | cut -c1-260
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H_mask2.log 2>&1; tail -1 gpurun_out/bench_H_mask2.log | cut -c1-260
