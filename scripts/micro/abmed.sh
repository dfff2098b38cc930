cd $GRAFT_REPO_ROOT
cp soc_project_stereo_matching_b200/lib/libsgm_b200.so /tmp/keep.so
for v in ${VARIANTS}; do
  cp scripts/micro/libs/$v.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
  echo "== $v"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"median_wavefront" -c 3 --csv python profiles/prof_frame.py 3 2>/dev/null | grep median | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
done
cp /tmp/keep.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
