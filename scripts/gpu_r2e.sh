# Round 2: lone-warp regime (one direction at a time: at most one warp per scheduler) - timings and ncu stall profiles.
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_e}
mkdir -p $O
for lay in 0 4 3; do
  for mask in 0x01 0x04 0x10; do
    echo "-- layout $lay dirmask $mask"
    SGM_B200_DEBUG_DIRMASK=$mask SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
  done
done
for cfg in "0 0x01" "0 0x04" "3 0x04"; do
  set -- $cfg
  SGM_B200_DEBUG_DIRMASK=$2 SGM_B200_DEBUG_LAYOUT=$1 python profiles/prof_frame.py 2 > $O/prof_plain_$1_$2.log 2>&1 && \
  SGM_B200_DEBUG_DIRMASK=$2 SGM_B200_DEBUG_LAYOUT=$1 ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate -s 1 -c 1 -o $O/lone_layout$1_mask$2 -f python profiles/prof_frame.py 2 > $O/ncu_$1_$2.log 2>&1; echo "ncu $1 $2 rc=$?"
done
