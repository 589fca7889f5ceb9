set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg"
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
# launch list of one eager step (our kernels live in namespace sivae)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sivae:: -c 1968 --csv --log-file gpurun_out/r02_7_launches.csv $B > gpurun_out/r02_7_ncu_l.log 2>&1; tail -1 gpurun_out/r02_7_ncu_l.log | cut -c1-200
timeout 500 $N -k regex:'k_conv_halo2<.*, 1>' -s 6 -c 6 -o gpurun_out/r02_7_prof_fwd_split $B > gpurun_out/r02_7_ncu_a.log 2>&1; tail -1 gpurun_out/r02_7_ncu_a.log | cut -c1-160
timeout 500 $N -k regex:'k_conv_halo2<.*, 2>' -s 0 -c 10 -o gpurun_out/r02_7_prof_dgrad16 $B > gpurun_out/r02_7_ncu_b.log 2>&1; tail -1 gpurun_out/r02_7_ncu_b.log | cut -c1-160
timeout 500 $N -k regex:'k_conv_wgrad_halo16|k_conv_wgrad_tc16' -s 0 -c 14 -o gpurun_out/r02_7_prof_wgrad16 $B > gpurun_out/r02_7_ncu_c.log 2>&1; tail -1 gpurun_out/r02_7_ncu_c.log | cut -c1-160
timeout 500 $N -k regex:'k_mse3_partial|k_mse3_final|k_adam|k_linear_|k_rowsep_|k_kl_reparam|k_latent_bwd|k_loss_seed|loss_finalize|k_split32|k_to_bf16|k_wg_reduce|k_splitk_reduce' -s 0 -c 60 -o gpurun_out/r02_7_prof_small $B > gpurun_out/r02_7_ncu_d.log 2>&1; tail -1 gpurun_out/r02_7_ncu_d.log | cut -c1-160
timeout 500 $N -k regex:'k_bn_act_fwd|k_bn_bwd_reduce|k_bn_bwd_apply|k_bn_bwd_small' -s 150 -c 40 -o gpurun_out/r02_7_prof_bn $B > gpurun_out/r02_7_ncu_e.log 2>&1; tail -1 gpurun_out/r02_7_ncu_e.log | cut -c1-160
timeout 500 $N -k regex:'k_conv_fwd_tc2|k_conv_halo<' -s 0 -c 16 -o gpurun_out/r02_7_prof_tc2_halo $B > gpurun_out/r02_7_ncu_f.log 2>&1; tail -1 gpurun_out/r02_7_ncu_f.log | cut -c1-160
ls -la gpurun_out/*.ncu-rep
for c in C M Bs; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg > gpurun_out/r02_7_bench_$c.json 2>/dev/null; head -c 250 gpurun_out/r02_7_bench_$c.json; echo; done
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x -k "split32 or bwd_bf16" > gpurun_out/r02_7_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_7_memcheck.log
