set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg"
N="ncu --set full --clock-control none --kernel-name-base demangled"
cap() { # name regex skip count
  timeout 400 $N -k regex:"$2" -s $3 -c $4 -o /tmp/$1 $B > gpurun_out/r02_9_ncu_$1.log 2>&1
  python profiles/summarize.py rep /tmp/$1.ncu-rep > gpurun_out/r02_9_prof_$1.md 2>gpurun_out/r02_9_sum_$1.err
  ls -la /tmp/$1.ncu-rep | awk '{print $5, $9}'; rm -f /tmp/$1.ncu-rep
}
cap fwd 'k_bn_act_fwd|k_conv_halo2<.*\(int\)1>' 16 8
cap bwd 'k_bn_bwd_reduce|k_bn_bwd_apply|k_conv_halo2<.*\(int\)2>|k_conv_wgrad_halo16' 0 12
cap misc 'k_mse3_partial|k_mse3_final|k_adam|k_linear_|k_kl_reparam|k_latent_bwd|k_loss_seed|loss_finalize' 0 44
cap rowsep 'k_rowsep_expand|k_rowsep_gather|k_rowsep_wg_reduce' 1 4
du -sh gpurun_out
