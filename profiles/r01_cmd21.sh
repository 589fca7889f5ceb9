set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
# 1. whole GPU suite (new: compensated 3xTF32 mode, full-size architecture parity)
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
# 2. headline bench (config H) with cpu_baseline
timeout 600 python bench.py --config H --steps 10 --warmup 3 --layers gpurun_out/layers_H.md > gpurun_out/bench_H.log 2>&1; tail -1 gpurun_out/bench_H.log | cut -c1-300
# 3. compensated mode at the same workload
SIVAE_CONV_BACKEND=3 timeout 300 python bench.py --config H --steps 3 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H_3x.md > gpurun_out/bench_H_3x.log 2>&1; tail -1 gpurun_out/bench_H_3x.log | cut -c1-300
# 4. the other BASELINE configs (parity cases; short timing for the record)
for c in C M Bs; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; tail -1 gpurun_out/bench_$c.log | cut -c1-200; done
# 5. ncu launch list of one eager step of the same command
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline"
SIVAE_CUDA_GRAPH=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 2100 -c 2008 --csv --log-file gpurun_out/launches_H.csv $B > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-120
# 6. ncu --set full of the dominant kernels (CTA-pair halo conv; then BN backward and halo wgrad)
SIVAE_CUDA_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo2 -s 40 -c 8 -o gpurun_out/prof_halo2 $B > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-120
SIVAE_CUDA_GRAPH=0 timeout 400 ncu --set full --clock-control none -k regex:"k_bn_bwd_apply|k_bn_bwd_reduce|k_conv_wgrad_halo" -s 20 -c 6 -o gpurun_out/prof_bwd $B > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-120
ls -la gpurun_out/
