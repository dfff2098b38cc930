cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/q97
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/q97/tests_gpu.log; tail -25 gpurun_out/q97/tests_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/q97/bench.json 2> gpurun_out/q97/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/q97/bench.json'))
print({k:j[k] for k in ('value','ms_per_step')}, 'agg_ms',j['roofline']['kernel_ms'])
print('e2e',j['e2e']['ms_per_step'],'batched',j['batched'])
print('census9x7',j['census9x7'])
PY
tail -3 gpurun_out/q97/bench.err
