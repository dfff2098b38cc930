cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 600 python scripts/prof_kernels.py c2 c1 c3 c5 2>/dev/null | cut -c1-420
