set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -6 gpurun_out/pytest_quick.log
timeout 200 python profiles/probe_conv_bw.py > gpurun_out/probe_bw10.log 2>&1; tail -10 gpurun_out/probe_bw10.log
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H21.md > gpurun_out/bench_H21.log 2>&1; tail -1 gpurun_out/bench_H21.log | cut -c1-200
