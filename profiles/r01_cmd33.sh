set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -4
timeout 300 python -m pytest tests/test_gpu_dist.py -q > gpurun_out/pytest_dist3.log 2>&1; tail -2 gpurun_out/pytest_dist3.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_H_n2_r01w.log 2> gpurun_out/bench_H_n2_r01w.err; tail -1 gpurun_out/bench_H_n2_r01w.log | cut -c1-300; tail -2 gpurun_out/bench_H_n2_r01w.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-loader-leg > gpurun_out/bench_H_n1_r01w.log 2>&1; tail -1 gpurun_out/bench_H_n1_r01w.log | cut -c1-200
