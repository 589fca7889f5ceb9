set -x
SIVAE_TC_2CTA=1 timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "fwd_dgrad and (128-3 or 256-3 or 320-3)" > gpurun_out/pytest_2cta.log 2>&1; tail -15 gpurun_out/pytest_2cta.log
SIVAE_TC_2CTA=1 PROBE_SHORT=1 timeout 120 python profiles/probe_conv_bw.py > gpurun_out/probe_2cta.log 2>&1; tail -4 gpurun_out/probe_2cta.log
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H18.md > gpurun_out/bench_H18.log 2>&1; tail -1 gpurun_out/bench_H18.log | cut -c1-200
