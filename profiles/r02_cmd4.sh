set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_4_dist.log
tail -30 gpurun_out/r02_4_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_4_bench2.json 2> gpurun_out/r02_4_bench2.err
head -c 1200 gpurun_out/r02_4_bench2.json; tail -5 gpurun_out/r02_4_bench2.err
SIVAE_DP_COMM=torch timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 --no-reuse-leg > gpurun_out/r02_4_bench2_torch.json 2> gpurun_out/r02_4_bench2_torch.err
head -c 600 gpurun_out/r02_4_bench2_torch.json; tail -3 gpurun_out/r02_4_bench2_torch.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg > gpurun_out/r02_4_bench1.json 2>/dev/null; head -c 400 gpurun_out/r02_4_bench1.json
