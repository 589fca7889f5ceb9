set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu5.log 2>&1; tail -4 gpurun_out/pytest_gpu5.log
timeout 300 python bench.py --config H --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H_reuse.log 2>&1; tail -1 gpurun_out/bench_H_reuse.log | cut -c1-200; grep -o '"decoder_pass_reuse": {[^}]*}' gpurun_out/bench_H_reuse.log
timeout 400 python tests/stock_torch_probe.py --config H --steps 4 --warmup 2 > gpurun_out/stock_torch_H.json 2> gpurun_out/stock_torch_H.err; cat gpurun_out/stock_torch_H.json; tail -3 gpurun_out/stock_torch_H.err
