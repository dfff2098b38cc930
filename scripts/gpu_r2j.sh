cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_j}
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -x 2>&1 | tail -3
SGM_B200_DEBUG_LAYOUT=3 timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x 2>&1 | tail -3
for lay in 3 0; do
  echo "-- layout $lay irregular only"
  SGM_B200_DEBUG_DIRMASK=0x100 SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
  for mask in 0x04 0xFC; do
    echo "-- layout $lay dirmask $mask NOIRR"
    SGM_B200_DEBUG_NOIRR=1 SGM_B200_DEBUG_DIRMASK=$mask SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
  done
  echo "-- layout $lay full"
  SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 c2p4 c1 c3 --no-e2e 2>/dev/null | cut -c1-140
done
