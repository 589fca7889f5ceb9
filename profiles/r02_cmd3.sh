set -x
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_report.jsonl
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_3_gpu_tests.log
tail -40 gpurun_out/r02_3_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r02_3_bench.json 2> gpurun_out/r02_3_bench.err
cat gpurun_out/r02_3_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','cpu_baseline','stock_pytorch_same_gpu','clocks')})
"
tail -5 gpurun_out/r02_3_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_3_ref.json 2> gpurun_out/r02_3_ref.err; cat gpurun_out/r02_3_ref.json; tail -3 gpurun_out/r02_3_ref.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
