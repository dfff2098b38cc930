# Round 2, first GPU call: smoke, the whole GPU suite (new full-size fixtures, drop-in run, stress), bench of the current
# tree, per-kernel timings, ncu --set full of K1 / K4 / K5.
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-r2_a}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -60 > $O/tests_gpu.log; tail -30 $O/tests_gpu.log
timeout 600 python scripts/prof_kernels.py > $O/kernels.jsonl 2> $O/kernels.err; echo "kernels rc=$?"; cat $O/kernels.jsonl; tail -3 $O/kernels.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
python profiles/prof_frame.py 2 > $O/prof_plain.log 2>&1 && for k in ${KERNELS:-sgm_census speckle_init speckle_merge speckle_count median_prepare median_wavefront}; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $O/full_$k -f python profiles/prof_frame.py 2 > $O/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
done
