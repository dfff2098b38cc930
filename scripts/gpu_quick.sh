cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/tests_gpu.log; tail -4 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench.json'))
    print({k:j[k] for k in ('value','ms_per_step')}, 'agg_ms',j['roofline']['kernel_ms'],'frac',round(j['roofline']['frac'],3),'frame_frac',round(j['roofline']['frame']['frac'],3))
    print('e2e',j['e2e']['ms_per_step'],'hot e2e',j['e2e']['hotpath_only']['ms_per_step'],'batched',j['batched'],'lat',j['latency_ms'], j['clocks'])
except Exception as e: print('ERR',e)
PY
tail -3 gpurun_out/bench.err
python profiles/prof_frame.py 3 > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv python profiles/prof_frame.py 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for row in csv.DictReader(lines):
    agg[row['Kernel Name'][:48]][row['Metric Name']].append(float(row['Metric Value'].replace(',','')))
for k,v in agg.items():
    t=v['gpu__time_duration.sum']; i=v['smsp__inst_executed.sum']
    print(f"{k:48s} n={len(t):3d} avg_us={sum(t)/len(t)/1e3:9.1f} inst={sum(i)/len(i)/1e6:8.2f}M")
PY
