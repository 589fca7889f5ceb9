set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fwd_dgrad" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
for i in 1 2; do
SIVAE_LIB_PATH=$PWD/profiles/_ab_prev.so SIVAE_TC_ADDEND=3 timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_prev$i.log 2>&1; tail -1 gpurun_out/bench_ab_prev$i.log | cut -c1-180
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_H26.md > gpurun_out/bench_ab_new$i.log 2>&1; tail -1 gpurun_out/bench_ab_new$i.log | cut -c1-180
done
PROBE_SHORT=1 timeout 200 python profiles/probe_conv_bw.py 2>&1 | tail -4
SIVAE_LIB_PATH=$PWD/profiles/_ab_prev.so SIVAE_TC_ADDEND=3 PROBE_SHORT=1 timeout 200 python profiles/probe_conv_bw.py 2>&1 | tail -4
