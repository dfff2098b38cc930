# ncu --set full of the aggregation kernel: the whole kernel and the column-like directions alone (results wrong with the mask).
cd $GRAFT_REPO_ROOT
O=gpurun_out/k2ncu; mkdir -p $O
python profiles/prof_frame.py 2 > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"sgm_aggregate" -s 1 -c 1 -o $O/all -f python profiles/prof_frame.py 2 > $O/ncu_all.log 2>&1; echo "rc=$?"
SGM_B200_DEBUG_DIRMASK=0xfc SGM_B200_DEBUG_NOIRR=1 ncu --set full --clock-control none -k regex:"sgm_aggregate" -s 1 -c 1 -o $O/cols -f python profiles/prof_frame.py 2 > $O/ncu_cols.log 2>&1; echo "rc=$?"
SGM_B200_DEBUG_DIRMASK=0x03 SGM_B200_DEBUG_NOIRR=1 ncu --set full --clock-control none -k regex:"sgm_aggregate" -s 1 -c 1 -o $O/hor -f python profiles/prof_frame.py 2 > $O/ncu_hor.log 2>&1; echo "rc=$?"
ls -la $O
