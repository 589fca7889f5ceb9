set -x
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 300 python bench.py --config H --steps 3 --warmup 2 --no-cpu-baseline --layers gpurun_out/layers_H7.md > gpurun_out/bench_H7.log 2>&1; tail -1 gpurun_out/bench_H7.log | cut -c1-400
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1900 -c 1900 --csv --log-file gpurun_out/launches_H2.csv python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1; tail -1 gpurun_out/ncu_bench2.log | cut -c1-200
