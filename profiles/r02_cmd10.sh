set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg --layers gpurun_out/r02_10_layers.md > gpurun_out/r02_10_bench.json 2> gpurun_out/r02_10_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_10_bench.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in ('fwd_split32','dgrad_bf16','wgrad_bf16','tf32_conv','wgrad_tf32','bn_fwd','bn_bwd'):
    print(k, r[k]['ms_per_step'], r[k].get('achieved'))
PY
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "bn_act" 2>&1 | tail -2
