cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 420 python -m pytest tests/test_gpu_step.py -q -m gpu -k "l1 or bce or conditional or auto_backend or celeb1024 or tiny_step_vs_reference_golden or graph_replay or inference_api" --durations=8 2>&1 | grep -v "it/s" | tail -60 > gpurun_out/r02_23_new_tests.log
tail -45 gpurun_out/r02_23_new_tests.log
cp gpurun_out/parity_report.jsonl gpurun_out/r02_23_parity_report.jsonl 2>/dev/null
timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "mse3 or kl_reparam" 2>&1 | tail -2
