cd $GRAFT_REPO_ROOT
python profiles/prof_frame.py 2 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sgm_aggregate|sgm_reduce_wta|median3|speckle_merge" -s 4 -c 4 -o gpurun_out/prof_full -f python profiles/prof_frame.py 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
