set -x
cd $GRAFT_REPO_ROOT
timeout 320 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_6_dist.log
tail -8 gpurun_out/r02_6_dist.log
for ov in 1 0; do
SIVAE_DP_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$ov bench.py --gpus 2 --steps 30 --warmup 3 --no-reuse-leg > gpurun_out/r02_6_bench2_ov$ov.json 2> gpurun_out/r02_6_bench2_ov$ov.err
echo "bench rc=$?"
grep '^{' gpurun_out/r02_6_bench2_ov$ov.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('overlap=$ov', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n1', d['value'], d['ms_per_step'], d['e2e']['value'])"
