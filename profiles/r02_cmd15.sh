cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_step.py -q -m gpu -k "resume or stale or graph_replay or reuse or train_driver or nan_loss" 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-leg --no-loader-leg --no-reuse-leg 2>/dev/null | head -c 250
