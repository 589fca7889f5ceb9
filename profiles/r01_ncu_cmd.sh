set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -q -m gpu -k "adam or tiny_step or fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1900 -c 1900 --csv --log-file gpurun_out/launches_H.csv python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_fwd_tc -s 150 -c 3 -o gpurun_out/prof_fwd python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fwd.log 2>&1; tail -2 gpurun_out/ncu_fwd.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_wgrad_tc -s 40 -c 2 -o gpurun_out/prof_wgrad python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_wgrad.log 2>&1; tail -2 gpurun_out/ncu_wgrad.log | cut -c1-200
timeout 400 ncu --set full --clock-control none -k regex:"k_bn_act_fwd|k_bn_bwd_apply|k_bn_stats_partial" -s 60 -c 4 -o gpurun_out/prof_bn python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bn.log 2>&1; tail -2 gpurun_out/ncu_bn.log | cut -c1-200
ls -la gpurun_out/
