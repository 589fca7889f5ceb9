cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py -q -m gpu -k "tracks" 2>&1 | grep -v "it/s" | tail -12
cat gpurun_out/parity_report.jsonl | tail -1
