# Round 2: ncu --set full of the aggregation kernel, round-1 lane-group layout (3) vs column-ILP layout (4, 0), plus timings.
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_c}
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x 2>&1 | tail -5
for lay in 0 4 3; do
  echo "== layout $lay"
  SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 c2p4 c3 --no-e2e > $O/kernels_layout$lay.jsonl 2> $O/kernels_layout$lay.err; echo "rc=$?"; cut -c1-420 $O/kernels_layout$lay.jsonl; tail -3 $O/kernels_layout$lay.err
done
for lay in 4 3; do
  SGM_B200_DEBUG_LAYOUT=$lay python profiles/prof_frame.py 2 > $O/prof_plain_$lay.log 2>&1 && \
  SGM_B200_DEBUG_LAYOUT=$lay ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate -s 1 -c 1 -o $O/full_agg_layout$lay -f python profiles/prof_frame.py 2 > $O/ncu_full_agg_$lay.log 2>&1; echo "ncu layout $lay rc=$?"
done
