cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_m}
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12
timeout 600 python scripts/prof_kernels.py c2 c2p4 c1 c3 c5 --no-e2e 2>/dev/null | cut -c1-330
python profiles/prof_frame.py 2 > $O/prof_plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sgm_reduce_wta -s 1 -c 1 -o $O/full_wta -f python profiles/prof_frame.py 2 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
