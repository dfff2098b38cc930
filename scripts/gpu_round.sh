# Full GPU round: smoke, parity tests, bench (both arms), launch list, ncu --set full of the hot-path kernels.
# Usage (from the repo root, here):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
cd $GRAFT_REPO_ROOT
TAG=${1:-rX}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > $O/tests_gpu.log; tail -4 $O/tests_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
if [ "${REFARM:-1}" = "1" ]; then timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cat $O/bench_reference.json; fi
python profiles/prof_frame.py 3 > $O/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 100 --csv --log-file $O/launches.csv python profiles/prof_frame.py 3 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"; cat $O/prof_plain.log
# launch list of the bench command itself (after it exited 0 without ncu above): the kernels' shares of a step must agree with the bench line
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --inflight 1 > $O/ncu_bench.log 2>&1; echo "ncu bench launches rc=$?"
for k in ${KERNELS:-sgm_aggregate sgm_reduce_wta}; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $O/full_$k -f python profiles/prof_frame.py 2 > $O/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
done
