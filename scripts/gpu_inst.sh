# instruction counts / durations of the aggregation kernel per direction mask and layout (ncu, metrics only)
cd $GRAFT_REPO_ROOT
export SGM_B200_DEBUG_NOIRR=${NOIRR:-1}
for lay in ${LAYOUTS:-0 1}; do
for m in ${MASKS:-0x01 0x04 0x10 0xff}; do
export SGM_B200_DEBUG_DIRMASK=$m SGM_B200_DEBUG_LAYOUT=$lay
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,launch__registers_per_thread --clock-control none -k regex:"sgm_aggregate" -s 1 -c 1 --csv python profiles/prof_frame.py 2 2>/dev/null | grep -E "sgm_aggregate" | awk -F'","' -v m=$m -v l=$lay '{gsub(/"/,"",$NF); printf "layout %s mask %s %-50s %s\n", l, m, $(NF-2), $NF}'
done; done
