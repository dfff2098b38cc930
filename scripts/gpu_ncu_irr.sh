cd $GRAFT_REPO_ROOT
export SGM_B200_DEBUG_DIRMASK=0x100
python profiles/prof_frame.py 2 > gpurun_out/prof_plain_irr.log 2>&1 && ncu --set full --clock-control none --cache-control none --import-source on -k regex:"sgm_aggregate" -s 1 -c 1 -o gpurun_out/prof_irr -f python profiles/prof_frame.py 2 > gpurun_out/ncu_irr.log 2>&1; echo "ncu rc=$?"
