cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-reuse-leg > gpurun_out/r02_16_bench8.json 2> gpurun_out/r02_16_bench8.err
echo "bench8 rc=$?"
grep '^{' gpurun_out/r02_16_bench8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n8', d['value'], d['ms_per_step'], d['e2e']['value'], d['launch_mode'][:80])"
tail -3 gpurun_out/r02_16_bench8.err
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n1', d['value'], d['ms_per_step'], d['e2e']['value'])"
