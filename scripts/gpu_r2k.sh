cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_k}
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
for lay in 0 4 3; do
  echo "-- layout $lay full"
  SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 c2p4 c1 c3 --no-e2e 2>/dev/null | cut -c1-140
done
for mask in 0x04 0x10 0xFC 0xFF; do
  echo "-- layout 0 dirmask $mask NOIRR"
  SGM_B200_DEBUG_NOIRR=1 SGM_B200_DEBUG_DIRMASK=$mask SGM_B200_DEBUG_LAYOUT=0 timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
done
echo "-- irregular only"; SGM_B200_DEBUG_DIRMASK=0x100 timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
SGM_B200_DEBUG_LAYOUT=0 python profiles/prof_frame.py 2 > $O/prof_plain_full.log 2>&1 && \
SGM_B200_DEBUG_LAYOUT=0 ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate -s 1 -c 1 -o $O/full_layout0 -f python profiles/prof_frame.py 2 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
