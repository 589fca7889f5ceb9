cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/jpeg_report.json
# 1. the opt-in nvJPEG path in its own process (a library crash must not take the suite with it)
timeout 120 python -m pytest tests/test_gpu_image.py -q -m gpu -k "gpu_jpeg" 2>&1 | grep -v "it/s" | tail -25 > gpurun_out/r02_25_jpeg_test.log
echo "jpeg rc=$?"; tail -12 gpurun_out/r02_25_jpeg_test.log; cat gpurun_out/jpeg_report.json 2>/dev/null; echo
# 2. the full GPU suite (the driver's command, minus the test above)
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_image.py::test_gpu_jpeg_decode_opt_in 2>&1 | grep -v "it/s" | tail -15 > gpurun_out/r02_25_gpu_tests.log
tail -6 gpurun_out/r02_25_gpu_tests.log
cp gpurun_out/parity_report.jsonl gpurun_out/r02_25_parity_report.jsonl 2>/dev/null
# 3. smoke
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# 4. the driver's bench command
timeout 420 python bench.py --layers gpurun_out/r02_25_layers_H.md > gpurun_out/r02_25_bench_H.json 2> gpurun_out/r02_25_bench_H.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_25_bench_H.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d.get('stock_pytorch_same_gpu',{}).get('ms_per_step'), d.get('cpu_baseline',{}).get('value'))
PY
tail -2 gpurun_out/r02_25_bench_H.err
