cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_q}
mkdir -p $O
python profiles/prof_frame.py 2 > $O/prof_plain.log 2>&1 && for k in median_wavefront speckle_merge; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $O/full_$k -f python profiles/prof_frame.py 2 > $O/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
done
