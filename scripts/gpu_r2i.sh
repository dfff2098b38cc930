cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_i}
mkdir -p $O
SGM_B200_DEBUG_LAYOUT=0 python profiles/prof_frame.py 2 > $O/prof_plain_full.log 2>&1 && \
SGM_B200_DEBUG_LAYOUT=0 ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate -s 1 -c 1 -o $O/full_layout0 -f python profiles/prof_frame.py 2 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
export SGM_B200_DEBUG_NOIRR=1
SGM_B200_DEBUG_DIRMASK=0x04 SGM_B200_DEBUG_LAYOUT=0 python profiles/prof_frame.py 2 > $O/prof_plain_lone.log 2>&1 && \
SGM_B200_DEBUG_DIRMASK=0x04 SGM_B200_DEBUG_LAYOUT=0 ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate -s 1 -c 1 -o $O/lone_layout0_mask0x04 -f python profiles/prof_frame.py 2 > $O/ncu_lone.log 2>&1; echo "ncu lone rc=$?"
