cd $GRAFT_REPO_ROOT
for lay in 8 0; do
  echo "-- layout $lay parity"; SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x 2>&1 | tail -2
  echo "-- layout $lay full"
  SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 c2p4 c1 c3 --no-e2e 2>/dev/null | cut -c1-140
done
for mask in 0x04 0xFC; do
  echo "-- layout 8 dirmask $mask NOIRR"
  SGM_B200_DEBUG_NOIRR=1 SGM_B200_DEBUG_DIRMASK=$mask SGM_B200_DEBUG_LAYOUT=8 timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
done
