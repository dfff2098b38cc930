# ncu --set full of selected kernels at C2.  Usage: gpurun -- 'KERNELS="a b" bash scripts/gpu_ncu.sh <tag>'
cd $GRAFT_REPO_ROOT
TAG=${1:-rX}
O=gpurun_out/$TAG
mkdir -p $O
python profiles/prof_frame.py 3 > $O/prof_plain.log 2>&1 || { cat $O/prof_plain.log; exit 1; }
cat $O/prof_plain.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 100 --csv --log-file $O/launches.csv python profiles/prof_frame.py 3 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for k in ${KERNELS:-sgm_aggregate sgm_reduce_wta median3}; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $O/full_$k -f python profiles/prof_frame.py 2 > $O/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
done
