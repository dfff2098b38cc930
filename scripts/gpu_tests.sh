# GPU parity tests only.  Usage: gpurun --timeout 900 -- 'bash scripts/gpu_tests.sh [pytest -k expression]'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/t
if [ -n "$1" ]; then
  timeout 800 python -m pytest tests -q -m gpu -x -k "$1" 2>&1 | tail -30 | tee gpurun_out/t/tests_gpu.log
else
  timeout 800 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 | tee gpurun_out/t/tests_gpu.log
fi
