set -x
SIVAE_TC_ADDEND=3 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_add3.log 2>&1; tail -6 gpurun_out/pytest_add3.log
SIVAE_TC_ADDEND=3 SIVAE_TC_2CTA=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_add3b.log 2>&1; tail -3 gpurun_out/pytest_add3b.log
SIVAE_TC_ADDEND=3 timeout 200 python profiles/probe_conv_bw.py > gpurun_out/probe_add3.log 2>&1; tail -10 gpurun_out/probe_add3.log
SIVAE_TC_ADDEND=3 timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H24.log 2>&1; tail -1 gpurun_out/bench_H24.log | cut -c1-200
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H24b.log 2>&1; tail -1 gpurun_out/bench_H24b.log | cut -c1-200
