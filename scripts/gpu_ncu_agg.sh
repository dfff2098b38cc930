cd $GRAFT_REPO_ROOT
python profiles/prof_frame.py 2 > gpurun_out/prof_plain_agg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sgm_aggregate" -s 1 -c 1 -o gpurun_out/prof_agg -f python profiles/prof_frame.py 2 > gpurun_out/ncu_agg.log 2>&1; echo "ncu rc=$?"
