set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu6.log 2>&1; tail -6 gpurun_out/pytest_gpu6.log
timeout 120 python profiles/probe_linear.py > gpurun_out/probe_linear.txt 2>&1; cat gpurun_out/probe_linear.txt
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H_r01t.log 2>&1; tail -1 gpurun_out/bench_H_r01t.log | cut -c1-230; grep -o '"image_loader": {.*' gpurun_out/bench_H_r01t.log | cut -c1-700
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_image_batch -s 2 -c 1 -o gpurun_out/prof_image2 -f python profiles/probe_image_kernel.py > gpurun_out/ncu_image2.log 2>&1; tail -2 gpurun_out/ncu_image2.log
