set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -x > gpurun_out/pytest_step.log 2>&1; tail -3 gpurun_out/pytest_step.log
timeout 300 python profiles/probe_conv_bw.py > gpurun_out/probe_bw.log 2>&1; cat gpurun_out/probe_bw.log | tail -12
timeout 300 python bench.py --config H --steps 3 --warmup 2 --no-cpu-baseline --layers gpurun_out/layers_H15.md > gpurun_out/bench_H15.log 2>&1; tail -1 gpurun_out/bench_H15.log | cut -c1-300
