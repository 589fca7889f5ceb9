set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_image.py -q -x > gpurun_out/pytest_gpu_image.log 2>&1; tail -15 gpurun_out/pytest_gpu_image.log
timeout 300 python bench.py --config H --steps 5 --warmup 3 --no-cpu-baseline --no-reuse-leg > gpurun_out/bench_H_loader.log 2>&1; grep -o '"image_loader": {.*' gpurun_out/bench_H_loader.log | cut -c1-1500
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_image_batch -s 2 -c 1 -o gpurun_out/prof_image -f python profiles/probe_image_kernel.py > gpurun_out/ncu_image.log 2>&1; tail -3 gpurun_out/ncu_image.log
