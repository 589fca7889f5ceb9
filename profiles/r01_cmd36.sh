set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu11.log 2>&1; tail -3 gpurun_out/pytest_gpu11.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; tail -1 gpurun_out/smoke2.log
( time timeout 900 python bench.py --layers gpurun_out/layers_H_r01y.md ) > gpurun_out/bench_default2.log 2> gpurun_out/bench_default2.err; tail -1 gpurun_out/bench_default2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['decoder_pass_reuse']['ms_per_step'], d['image_loader']['value'], d['cpu_baseline']['value'], d['clocks'])"; tail -3 gpurun_out/bench_default2.err
