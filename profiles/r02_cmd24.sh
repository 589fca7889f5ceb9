cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_image.py -q -m gpu 2>&1 | grep -v "it/s" | tail -40 > gpurun_out/r02_24_image_tests.log
tail -40 gpurun_out/r02_24_image_tests.log
timeout 200 python -m pytest tests/test_gpu_step.py -q -m gpu -k "tiny_step_l1_bce" 2>&1 | tail -3
timeout 100 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_image.py -q -m gpu -k "window_and_u8 or load_image_branches" 2>&1 | tail -6 > gpurun_out/r02_24_memcheck_image.log
tail -6 gpurun_out/r02_24_memcheck_image.log
