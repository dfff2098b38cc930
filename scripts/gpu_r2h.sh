cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_h}
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x 2>&1 | tail -3
for lay in 0; do
  for mask in 0x04 0x10 0xFC 0xFF; do
    echo "-- layout $lay dirmask $mask NOIRR"
    SGM_B200_DEBUG_NOIRR=1 SGM_B200_DEBUG_DIRMASK=$mask SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
  done
  echo "-- layout $lay full"
  SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 c2p4 c3 --no-e2e 2>/dev/null | cut -c1-140
done
