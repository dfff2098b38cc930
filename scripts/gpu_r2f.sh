# Round 2: is the floor of the aggregation kernel the four irregular-path warps?  One direction at a time, with and without them.
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_f}
mkdir -p $O
for lay in 0 3; do
  for mask in 0x01 0x04 0x10 0xFF; do
    echo "-- layout $lay dirmask $mask NOIRR"
    SGM_B200_DEBUG_NOIRR=1 SGM_B200_DEBUG_DIRMASK=$mask SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
  done
  echo "-- layout $lay dirmask 0x00 (irregular warps only)"
  SGM_B200_DEBUG_DIRMASK=0x100 SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
done
