cd $GRAFT_REPO_ROOT
for hh in 32 64 128 375; do
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:median_wavefront -c 4 --csv python profiles/prof_frame.py 2 1242x${hh}x16 2>/dev/null | grep median_wavefront | awk -F'","' -v h=$hh '{print h, $NF}'
done
