set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_boundary.py -q -m gpu -k "resume or fid or reference_main or bootstrap_vae or train_driver" 2>&1 | tail -25 > gpurun_out/r02_12_tests.log
tail -25 gpurun_out/r02_12_tests.log
