set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
SIVAE_LIB_PATH=$PWD/profiles/_ab_prev.so SIVAE_TC_ADDEND=3 timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_prev.md > gpurun_out/bench_ab_prev1.log 2>&1; tail -1 gpurun_out/bench_ab_prev1.log | cut -c1-180
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_new.md > gpurun_out/bench_ab_new1.log 2>&1; tail -1 gpurun_out/bench_ab_new1.log | cut -c1-180
SIVAE_LIB_PATH=$PWD/profiles/_ab_prev.so SIVAE_TC_ADDEND=1 timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_prev_m1.md > gpurun_out/bench_ab_prev2.log 2>&1; tail -1 gpurun_out/bench_ab_prev2.log | cut -c1-180
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_new2.log 2>&1; tail -1 gpurun_out/bench_ab_new2.log | cut -c1-180
