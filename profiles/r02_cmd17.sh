set -x
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_report.jsonl
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_17_gpu_tests.log
tail -8 gpurun_out/r02_17_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_17_bench.json 2> gpurun_out/r02_17_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_17_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['stock_pytorch_same_gpu']['ms_per_step'], d['decoder_pass_reuse']['ms_per_step'])
r=d['roofline']
print(r['frac'], r['tensor_pipe_issued_frac'], {k:(r[k] or {}).get('ms_per_step') for k in ('fwd_split32','dgrad_bf16','wgrad_bf16','tf32_conv','wgrad_tf32','bn_fwd','bn_bwd')}, r['conv_stack'])
PY
tail -3 gpurun_out/r02_17_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_17_ref.json 2>/dev/null; head -c 200 gpurun_out/r02_17_ref.json
