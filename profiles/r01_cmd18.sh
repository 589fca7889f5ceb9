set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 1 -c 1 -o gpurun_out/prof_stem python profiles/probe_one_conv.py 32 256 256 3 64 5 0 > gpurun_out/ncu_stem.log 2>&1; tail -1 gpurun_out/ncu_stem.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_fwd_tc2 -s 1 -c 1 -o gpurun_out/prof_1x1 python profiles/probe_one_conv.py 32 128 128 64 128 1 2 > gpurun_out/ncu_1x1.log 2>&1; tail -1 gpurun_out/ncu_1x1.log
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H25.log 2>&1; tail -1 gpurun_out/bench_H25.log | cut -c1-200
