set -x
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv" 2>&1 | tail -25 > gpurun_out/r02_2_kernels.log
tail -25 gpurun_out/r02_2_kernels.log
timeout 1500 python -m pytest tests/test_gpu_step.py -q -x 2>&1 | tail -30 > gpurun_out/r02_2_step.log
tail -30 gpurun_out/r02_2_step.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-loader-leg --layers gpurun_out/r02_2_layers.md > gpurun_out/r02_2_bench.json 2> gpurun_out/r02_2_bench.err
cat gpurun_out/r02_2_bench.json | head -c 1500; tail -5 gpurun_out/r02_2_bench.err
