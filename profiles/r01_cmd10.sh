set -x
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_all.log 2>&1; tail -4 gpurun_out/pytest_all.log
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H17.md > gpurun_out/bench_H17.log 2>&1; tail -1 gpurun_out/bench_H17.log | cut -c1-400
