cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | grep -v "it/s" | tail -6 > gpurun_out/r02_26_dist_tests.log
tail -6 gpurun_out/r02_26_dist_tests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --no-reuse-leg --no-loader-leg > gpurun_out/r02_26_bench_n2.json 2> gpurun_out/r02_26_bench_n2.err
echo "bench2 rc=$?"
grep '^{' gpurun_out/r02_26_bench_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n2', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d.get('comm'))"
tail -2 gpurun_out/r02_26_bench_n2.err
