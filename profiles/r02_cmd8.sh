set -x
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/*.ncu-rep
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg"
N="ncu --set full --clock-control none --kernel-name-base demangled"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sivae:: -c 1968 --csv --log-file gpurun_out/r02_8_launches.csv $B > gpurun_out/r02_8_ncu_l.log 2>&1
python profiles/summarize.py launches gpurun_out/r02_8_launches.csv > gpurun_out/r02_8_launches_H.md; gzip -f gpurun_out/r02_8_launches.csv; tail -3 gpurun_out/r02_8_launches_H.md
cap() { # name regex skip count
  timeout 400 $N -k regex:"$2" -s $3 -c $4 -o /tmp/$1 $B > gpurun_out/r02_8_ncu_$1.log 2>&1
  python profiles/summarize.py rep /tmp/$1.ncu-rep > gpurun_out/r02_8_prof_$1.md 2>gpurun_out/r02_8_sum_$1.err
  ls -la /tmp/$1.ncu-rep | awk '{print $5, $9}'; rm -f /tmp/$1.ncu-rep
}
cap fwd_split 'k_conv_halo2<.*, 1>' 6 4
cap dgrad16 'k_conv_halo2<.*, 2>' 0 6
cap wgrad16 'k_conv_wgrad_halo16|k_conv_wgrad_tc16' 0 10
cap small 'k_mse3_partial|k_mse3_final|k_adam|k_linear_|k_rowsep_|k_kl_reparam|k_latent_bwd|k_loss_seed|loss_finalize|k_split32|k_to_bf16|k_wg_reduce|k_splitk_reduce|k_colsum' 0 48
cap bn 'k_bn_act_fwd|k_bn_bwd_reduce|k_bn_bwd_apply|k_bn_bwd_small|k_bn_stats_finalize' 170 14
cap tc2_halo 'k_conv_fwd_tc2|k_conv_halo<' 0 10
du -sh gpurun_out
