set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -4 gpurun_out/pytest_quick.log
SIVAE_TC_2CTA=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_2cta2.log 2>&1; tail -4 gpurun_out/pytest_2cta2.log
timeout 200 python profiles/probe_conv_bw.py > gpurun_out/probe_a.log 2>&1; tail -10 gpurun_out/probe_a.log
SIVAE_TC_2CTA=2 PROBE_SHORT=1 timeout 200 python profiles/probe_conv_bw.py > gpurun_out/probe_b.log 2>&1; tail -4 gpurun_out/probe_b.log
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H22.md > gpurun_out/bench_H22.log 2>&1; tail -1 gpurun_out/bench_H22.log | cut -c1-200
SIVAE_TC_2CTA=2 timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H22b.log 2>&1; tail -1 gpurun_out/bench_H22b.log | cut -c1-200
