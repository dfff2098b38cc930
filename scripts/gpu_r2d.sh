# Round 2: column-ILP layout with the single-buffer prefetch: parity, timings per layout and per direction family, ncu.
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_d}
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x 2>&1 | tail -5
for lay in 0 4 3; do
  echo "== layout $lay"
  SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 c2p4 c3 --no-e2e > $O/kernels_layout$lay.jsonl 2> $O/kernels_layout$lay.err; echo "rc=$?"; cut -c1-200 $O/kernels_layout$lay.jsonl; tail -3 $O/kernels_layout$lay.err
  for mask in 0x03 0xFC 0x0C 0xF0; do
    echo "-- layout $lay dirmask $mask"
    SGM_B200_DEBUG_DIRMASK=$mask SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 --no-e2e 2>/dev/null | cut -c1-140
  done
done
for lay in 0 4; do
  SGM_B200_DEBUG_LAYOUT=$lay python profiles/prof_frame.py 2 > $O/prof_plain_$lay.log 2>&1 && \
  SGM_B200_DEBUG_LAYOUT=$lay ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate -s 1 -c 1 -o $O/full_agg_layout$lay -f python profiles/prof_frame.py 2 > $O/ncu_full_agg_$lay.log 2>&1; echo "ncu layout $lay rc=$?"
done
