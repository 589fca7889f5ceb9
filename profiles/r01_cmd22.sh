set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu2.log 2>&1; tail -4 gpurun_out/pytest_gpu2.log
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H_fast.md > gpurun_out/bench_H_fast.log 2>&1; tail -1 gpurun_out/bench_H_fast.log | cut -c1-260
SIVAE_TC_FASTEPI=0 timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H_slow.md > gpurun_out/bench_H_slow.log 2>&1; tail -1 gpurun_out/bench_H_slow.log | cut -c1-260
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H_fast2.log 2>&1; tail -1 gpurun_out/bench_H_fast2.log | cut -c1-260
for c in C M Bs; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_$c.log 2>&1; tail -1 gpurun_out/bench2_$c.log | cut -c1-200; done
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline"
SIVAE_CUDA_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo2 -s 40 -c 8 -o gpurun_out/prof_halo2_fast $B > gpurun_out/ncu_a2.log 2>&1; tail -1 gpurun_out/ncu_a2.log | cut -c1-120
SIVAE_CUDA_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_fwd_tc2 -s 30 -c 6 -o gpurun_out/prof_tc2_fast $B > gpurun_out/ncu_b2.log 2>&1; tail -1 gpurun_out/ncu_b2.log | cut -c1-120
