set -x
SIVAE_TC_2CTA=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_2cta.log 2>&1; tail -15 gpurun_out/pytest_2cta.log
SIVAE_TC_2CTA=1 timeout 200 python profiles/probe_conv_bw.py > gpurun_out/probe_2cta.log 2>&1; tail -8 gpurun_out/probe_2cta.log
SIVAE_TC_2CTA=1 timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H19.md > gpurun_out/bench_H19.log 2>&1; tail -1 gpurun_out/bench_H19.log | cut -c1-200
