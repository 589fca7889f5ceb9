set -x
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_all.log 2>&1; tail -4 gpurun_out/pytest_all.log
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H23.md > gpurun_out/bench_H23.log 2>&1; tail -1 gpurun_out/bench_H23.log | cut -c1-200
SIVAE_TC_NOFENCE=1 PROBE_SHORT=1 timeout 200 python profiles/probe_conv_bw.py > gpurun_out/probe_nofence.log 2>&1; tail -4 gpurun_out/probe_nofence.log
PROBE_SHORT=1 timeout 200 python profiles/probe_conv_bw.py > gpurun_out/probe_fence.log 2>&1; tail -4 gpurun_out/probe_fence.log
