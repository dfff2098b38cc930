cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_l}
mkdir -p $O
for lay in 0 5 6 7; do
  echo "-- layout $lay parity"; SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x 2>&1 | tail -2
  echo "-- layout $lay full"
  SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 c2p4 c1 c3 --no-e2e 2>/dev/null | cut -c1-140
done
