# Round-2 capture: smoke, the whole GPU suite, per-kernel timings, bench (both arms), launch list of the bench command,
# ncu --set full of the kernels named in $KERNELS (at most two reports per call: the merged output is capped at 64 MiB).
# Usage (from the repo root, here):  gpurun --timeout 1800 -- 'bash scripts/gpu_round2.sh <tag>'
cd $GRAFT_REPO_ROOT
TAG=${1:-r2_x}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1200 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -15 > $O/tests_gpu.log; tail -10 $O/tests_gpu.log
timeout 600 python scripts/prof_kernels.py c2 c2p4 c1 c3 c5 > $O/kernels.jsonl 2> $O/kernels.err; echo "kernels rc=$?"; cut -c1-400 $O/kernels.jsonl; tail -3 $O/kernels.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
if [ "${REFARM:-1}" = "1" ]; then timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cat $O/bench_reference.json; fi
if [ -n "${KERNELS:-}" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --inflight 1 --replays 2 --pool-pairs 0 > $O/ncu_bench.log 2>&1; echo "ncu bench launches rc=$?"
  python profiles/prof_frame.py 2 > $O/prof_plain.log 2>&1 && for k in $KERNELS; do
    ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $O/full_$k -f python profiles/prof_frame.py 2 > $O/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
  done
fi
