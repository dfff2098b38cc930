# Round-2 companion of gpu_round2.sh: refresh the bench line (reads the updated profiles/traffic.json), then one
# `ncu --set full` capture per small kernel; the reports are summarised ON the box (scripts/ncu_summary.py) and only the
# text summaries travel back (six reports would exceed the 64 MiB merge cap).
# Usage: gpurun --timeout 1500 -- 'KERNELS="sgm_census speckle_init ..." bash scripts/gpu_ncu_small.sh <tag>'
cd $GRAFT_REPO_ROOT
TAG=${1:-r2_x}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-300 $O/bench.json; tail -5 $O/bench.err
python profiles/prof_frame.py 2 > $O/prof_plain.log 2>&1 && for k in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o /tmp/full_$k -f python profiles/prof_frame.py 2 > $O/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
  python scripts/ncu_summary.py /tmp/full_$k.ncu-rep > $O/ncu_full_$k.txt 2>> $O/ncu_full_$k.log; wc -l $O/ncu_full_$k.txt
done
