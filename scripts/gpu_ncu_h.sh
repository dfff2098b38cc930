cd $GRAFT_REPO_ROOT
export SGM_B200_DEBUG_DIRMASK=0x01
python profiles/prof_frame.py 2 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sgm_aggregate" -s 1 -c 1 -o gpurun_out/prof_h -f python profiles/prof_frame.py 2 > gpurun_out/ncu_h.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_h.log
