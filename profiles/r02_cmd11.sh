set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_boundary.py -x -q -m gpu -k "resume or fid or reference_main or bootstrap_vae or train_driver" 2>&1 | tail -15 > gpurun_out/r02_11_tests.log
tail -15 gpurun_out/r02_11_tests.log
for c in C M Bs; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg > gpurun_out/r02_11_bench_$c.json 2>/dev/null; head -c 220 gpurun_out/r02_11_bench_$c.json; echo; done
