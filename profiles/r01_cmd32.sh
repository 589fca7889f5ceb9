set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
# 1. whole GPU suite + smoke
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu8.log 2>&1; tail -3 gpurun_out/pytest_gpu8.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
# 2. the default bench line (cpu_baseline, reuse leg, loader leg) and the reference arm, as the driver runs them
( time timeout 900 python bench.py --layers gpurun_out/layers_H_r01w.md ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c1-250; tail -4 gpurun_out/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.log | cut -c1-700; tail -4 gpurun_out/bench_reference.err | head -2
# 3. memcheck of the new kernels (image batch assembly, fc)
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_image.py tests/test_gpu_kernels.py -q -x -k "golden_vectors or (matches_oracle and 3 and 257) or (matches_oracle and 1 and 100) or downscale or linear_fwd_dgrad" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
# 4. ncu launch list of one eager step of the bench command
B="python bench.py --config H --steps 1 --warmup 1 --no-cpu-baseline --no-reuse-leg --no-loader-leg"
SIVAE_CUDA_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2100 -c 2008 --csv --log-file gpurun_out/launches_H_r01w.csv $B > gpurun_out/ncu_launches_r01w.log 2>&1; tail -1 gpurun_out/ncu_launches_r01w.log | cut -c1-120
