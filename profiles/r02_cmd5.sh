set -x
cd $GRAFT_REPO_ROOT
timeout 320 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_5_dist.log
tail -15 gpurun_out/r02_5_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-reuse-leg > gpurun_out/r02_5_bench2.json 2> gpurun_out/r02_5_bench2.err
echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02_5_bench2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launch_mode'])"
tail -3 gpurun_out/r02_5_bench2.err
