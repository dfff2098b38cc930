cd $GRAFT_REPO_ROOT
cp soc_project_stereo_matching_b200/lib/libsgm_b200.so /tmp/keep.so
cp scripts/micro/libsgm_dbg.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
SGM_B200_NO_GRAPH=1 python profiles/prof_frame.py 2 1242x375x16 2>&1 | grep "median g=" | tail -12
cp /tmp/keep.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
