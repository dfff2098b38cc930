cd $GRAFT_REPO_ROOT
export SGM_B200_DEBUG_NOIRR=1
for m in ${MASKS:-0x04}; do
export SGM_B200_DEBUG_DIRMASK=$m
python profiles/prof_frame.py 2 > gpurun_out/prof_plain_$m.log 2>&1 && ncu --set full --clock-control none --cache-control none --import-source on -k regex:"sgm_aggregate" -s 1 -c 1 -o gpurun_out/prof_$m -f python profiles/prof_frame.py 2 > gpurun_out/ncu_$m.log 2>&1; echo "ncu rc=$?"
done
