cd $GRAFT_REPO_ROOT
timeout 300 python profiles/tools/diag_resume.py 2>&1 | grep -v "it/s" | tail -40
