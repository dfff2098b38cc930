cd $GRAFT_REPO_ROOT
cp soc_project_stereo_matching_b200/lib/libsgm_b200.so /tmp/keep.so
for v in ${VARIANTS:-cur norot nowraph nowrapv imad allold}; do
  cp scripts/micro/libs/$v.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
  echo "== $v"; MASKS="${MASKS:-0xff}" python profiles/prof_dirs.py
done
cp /tmp/keep.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
