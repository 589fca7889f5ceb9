cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_boundary.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg --layers gpurun_out/r02_20_layers.md > gpurun_out/r02_20_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_20_bench.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], d['clocks'])
print({k:(r[k] or {}).get('ms_per_step') for k in ('fwd_split32','dgrad_bf16','wgrad_bf16','tf32_conv','wgrad_tf32','bn_fwd','bn_bwd')})
PY
grep " 3 | 64 | 5 \| 64 | 3 | 5 " gpurun_out/r02_20_layers.md
