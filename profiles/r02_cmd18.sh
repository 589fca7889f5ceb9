cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-loader-leg --no-reuse-leg > gpurun_out/r02_18_bench.json 2> gpurun_out/r02_18_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_18_bench.json'))
print(d['value'], d['ms_per_step'], d['stock_pytorch_same_gpu'])
PY
tail -3 gpurun_out/r02_18_bench.err
