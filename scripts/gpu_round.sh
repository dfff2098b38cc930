cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/tests_gpu.log; tail -5 gpurun_out/tests_gpu.log
nproc; free -g | head -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python profiles/prof_frame.py 3 > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python profiles/prof_frame.py 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/prof_plain.log
