cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -k "resume" 2>&1 | grep -v "^  \|it/s\|^        (\|^      (\|^    (" | grep -B2 -A25 "def test_resume\|Error" | head -120 > gpurun_out/r02_13.log
cat gpurun_out/r02_13.log | tail -80
