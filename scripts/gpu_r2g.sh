# Round 2: ncu stall profiles of the lone-warp regime (one direction, no irregular warps).
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_g}
mkdir -p $O
export SGM_B200_DEBUG_NOIRR=1
for cfg in "0 0x04" "3 0x04"; do
  set -- $cfg
  SGM_B200_DEBUG_DIRMASK=$2 SGM_B200_DEBUG_LAYOUT=$1 python profiles/prof_frame.py 2 > $O/prof_plain_$1_$2.log 2>&1 && \
  SGM_B200_DEBUG_DIRMASK=$2 SGM_B200_DEBUG_LAYOUT=$1 ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate -s 1 -c 1 -o $O/lone_layout$1_mask$2 -f python profiles/prof_frame.py 2 > $O/ncu_$1_$2.log 2>&1; echo "ncu $1 $2 rc=$?"
done
ls -la $O
