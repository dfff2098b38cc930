cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python scripts/prof_kernels.py c2 c2p4 c1 c3 --no-e2e 2>/dev/null | cut -c1-200
