set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -6
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -x > gpurun_out/pytest_dist.log 2>&1; tail -15 gpurun_out/pytest_dist.log
P=29533
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_H_n2_eager.log 2>&1; tail -1 gpurun_out/bench_H_n2_eager.log | cut -c1-300
SIVAE_CUDA_GRAPH=2 timeout 300 $TR --master-port $((P+1)) bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_H_n2_graph.log 2>&1; tail -1 gpurun_out/bench_H_n2_graph.log | cut -c1-300
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_H_n1.log 2>&1; tail -1 gpurun_out/bench_H_n1.log | cut -c1-300
timeout 200 $TR --master-port $((P+2)) bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-budget 5 --config C > gpurun_out/bench_ref_n2.log 2>&1; tail -2 gpurun_out/bench_ref_n2.log | cut -c1-300
