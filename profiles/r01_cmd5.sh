set -x
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -x > gpurun_out/pytest_step.log 2>&1; tail -3 gpurun_out/pytest_step.log
timeout 300 python bench.py --config H --steps 3 --warmup 2 --no-cpu-baseline --layers gpurun_out/layers_H12.md > gpurun_out/bench_H12.log 2>&1; tail -1 gpurun_out/bench_H12.log | cut -c1-300
timeout 300 python bench.py --config H --steps 3 --warmup 2 --no-cpu-baseline --graph > gpurun_out/bench_H12g.log 2>&1; tail -1 gpurun_out/bench_H12g.log | grep -o '"graph[^,]*,'
