cd $GRAFT_REPO_ROOT
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 3 --no-reuse-leg > gpurun_out/r02_22_bench4.json 2> gpurun_out/r02_22_bench4.err
echo "bench4 rc=$?"
grep '^{' gpurun_out/r02_22_bench4.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n4', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n1', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
