# Quick GPU check: parity tests, a short bench without the CPU baseline, launch list.
cd $GRAFT_REPO_ROOT
TAG=${1:-q}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > $O/tests_gpu.log; tail -6 $O/tests_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python - $O <<'PY'
import json, sys
try:
    j=json.load(open(sys.argv[1]+'/bench.json'))
    print({k:j[k] for k in ('value','ms_per_step')}, 'agg_ms',j['roofline']['kernel_ms'],'frac',round(j['roofline']['frac'],3))
    print('e2e',j['e2e']['ms_per_step'],'hot e2e',j['e2e']['hotpath_only']['ms_per_step'],'batched',j['batched'],'lat',j['latency_ms'], j['clocks'])
except Exception as e: print('ERR',e)
PY
tail -3 $O/bench.err
python profiles/prof_frame.py 3 > $O/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 100 --csv --log-file $O/launches.csv python profiles/prof_frame.py 3 > $O/ncu_launch.log 2>&1; echo "ncu rc=$?"; cat $O/prof_plain.log
python - $O <<'PY'
import csv, collections, sys
lines=[l for l in open(sys.argv[1]+'/launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for row in csv.DictReader(lines):
    agg[row['Kernel Name'][:48]][row['Metric Name']].append(float(row['Metric Value'].replace(',','')))
for k,v in agg.items():
    t=v['gpu__time_duration.sum']; i=v['smsp__inst_executed.sum']
    print(f"{k:48s} n={len(t):3d} avg_us={sum(t)/len(t)/1e3:9.1f} inst={sum(i)/len(i)/1e6:8.2f}M")
PY
