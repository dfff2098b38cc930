cd $GRAFT_REPO_ROOT
cp soc_project_stereo_matching_b200/lib/libsgm_b200.so /tmp/keep.so
cp scripts/micro/libsgm_dbg.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
python profiles/prof_frame.py 1 1242x32x16 2>&1 | tail -3
python profiles/prof_frame.py 1 1242x375x16 2>&1 | tail -14
cp /tmp/keep.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
