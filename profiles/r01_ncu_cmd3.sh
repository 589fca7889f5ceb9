set -x
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_conv_halo<64, 2" -s 20 -c 2 -o gpurun_out/prof_halo64 $B > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_conv_halo<128, 2" -s 20 -c 2 -o gpurun_out/prof_halo128 $B > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_narrow_in_fwd|k_narrow_corr<" -s 4 -c 2 -o gpurun_out/prof_narrow $B > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_conv_halo<32" -s 2 -c 1 -o gpurun_out/prof_halo5x5 $B > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_conv_wgrad_halo<128" -s 10 -c 2 -o gpurun_out/prof_wgradhalo $B > gpurun_out/ncu_e.log 2>&1; tail -1 gpurun_out/ncu_e.log | cut -c1-120
ls -la gpurun_out/*.ncu-rep
