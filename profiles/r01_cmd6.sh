set -x
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -x > gpurun_out/pytest_step.log 2>&1; tail -3 gpurun_out/pytest_step.log
timeout 300 python bench.py --config H --steps 3 --warmup 2 --no-cpu-baseline --layers gpurun_out/layers_H13.md > gpurun_out/bench_H13.log 2>&1; tail -1 gpurun_out/bench_H13.log | cut -c1-300
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_fwd_tc2 -s 30 -c 8 -o gpurun_out/prof_v2 $B > gpurun_out/ncu_v2.log 2>&1; tail -1 gpurun_out/ncu_v2.log | cut -c1-120
