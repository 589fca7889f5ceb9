set -x
mkdir -p gpurun_out
timeout 220 python -m pytest tests/test_gpu_dist.py -q -m gpu -x > gpurun_out/pytest_dist2.log 2>&1; tail -4 gpurun_out/pytest_dist2.log | cut -c1-1500
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_H_n2_seg.log 2>&1; tail -1 gpurun_out/bench_H_n2_seg.log | cut -c1-300
