set -x
timeout 120 python profiles/probe_linear_err.py > gpurun_out/probe_linear_err.txt 2>&1; cat gpurun_out/probe_linear_err.txt
timeout 600 python -m pytest tests/test_gpu_step.py -q > gpurun_out/pytest_gpu_step.log 2>&1; tail -5 gpurun_out/pytest_gpu_step.log
