set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -2 gpurun_out/pytest_quick.log
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 python bench.py --config H --steps 3 --warmup 2 --no-cpu-baseline --layers gpurun_out/layers_H9.md > gpurun_out/bench_H9.log 2>&1; tail -1 gpurun_out/bench_H9.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 200 -c 14 -o gpurun_out/prof_halo $B > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_narrow_corr -s 7 -c 1 -o gpurun_out/prof_corr $B > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_wgrad_halo -s 70 -c 4 -o gpurun_out/prof_wgradhalo $B > gpurun_out/ncu_e.log 2>&1; tail -1 gpurun_out/ncu_e.log | cut -c1-120
ls -la gpurun_out/*.ncu-rep
