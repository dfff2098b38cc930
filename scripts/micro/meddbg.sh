cd $GRAFT_REPO_ROOT
cp scripts/micro/libsgm_dbg.so soc_project_stereo_matching_b200/lib/libsgm_b200.so
python profiles/prof_frame.py 1 1242x64x16 2>&1 | tail -4
python profiles/prof_frame.py 1 1242x375x16 2>&1 | tail -14
