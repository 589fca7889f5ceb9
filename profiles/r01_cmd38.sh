set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_image.py -q -k "train_driver or nan_loss or prefetcher or loader_end_to_end" > gpurun_out/pytest_gpu12.log 2>&1; tail -2 gpurun_out/pytest_gpu12.log
timeout 900 python bench.py > gpurun_out/bench_default3.log 2> gpurun_out/bench_default3.err; tail -1 gpurun_out/bench_default3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['decoder_pass_reuse']['ms_per_step'], d['cpu_baseline']['value'])"; tail -2 gpurun_out/bench_default3.err
