# Round 2: correctness of the column-ILP aggregation layout + timings of the layouts side by side.
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_b}
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 1200 python -m pytest tests -q -m gpu -x --durations=5 2>&1 | tail -40 > $O/tests_gpu.log; tail -25 $O/tests_gpu.log
for lay in 0 4 3; do
  echo "== layout $lay"
  SGM_B200_DEBUG_LAYOUT=$lay timeout 600 python scripts/prof_kernels.py c2 c2p4 c1 c3 --no-e2e > $O/kernels_layout$lay.jsonl 2> $O/kernels_layout$lay.err; echo "rc=$?"; cat $O/kernels_layout$lay.jsonl; tail -3 $O/kernels_layout$lay.err
done
