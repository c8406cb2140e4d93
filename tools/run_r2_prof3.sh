# round 2, final kernels: launch list of one step, DRAM traffic of one factorisation, ncu --set full of the leaf and of the
# TRSV sweep, per-panel trace.  Captures go to /tmp, only text summaries come back.
export TMPD=/tmp/lgpprof; mkdir -p $TMPD
python tools/one_step.py 20000 > gpurun_out/plain_step.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_step_r2b.csv python tools/one_step.py 20000 > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_step_r2b.csv > gpurun_out/launches_step_r2b_summary.txt 2>&1
gzip -f gpurun_out/launches_step_r2b.csv
python tools/prof_chol.py 20000 factor > gpurun_out/plain_chol.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $TMPD/traffic.csv python tools/prof_chol.py 20000 factor > gpurun_out/ncu_traffic.log 2>&1
python tools/summarize_traffic.py $TMPD/traffic.csv 'gram_' > gpurun_out/traffic_chol20k_r2.txt 2>&1
python tools/bench_leaf.py > gpurun_out/plain_leaf.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:potrf_leaf3 -s 2 -c 1 -o $TMPD/prof_leaf3 python tools/bench_leaf.py > gpurun_out/ncu_leaf3.log 2>&1
ncu -i $TMPD/prof_leaf3.ncu-rep --page details > gpurun_out/potrf_leaf3_details.txt 2>&1
ncu -i $TMPD/prof_leaf3.ncu-rep --page raw --csv > gpurun_out/potrf_leaf3_raw.csv 2>&1
python tools/time_solve.py 20000 > gpurun_out/plain_solve.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:trsv_sweep -s 2 -c 2 -o $TMPD/prof_trsv python tools/time_solve.py 20000 > gpurun_out/ncu_trsv.log 2>&1
ncu -i $TMPD/prof_trsv.ncu-rep --page details > gpurun_out/trsv_sweep_details.txt 2>&1
ncu -i $TMPD/prof_trsv.ncu-rep --page raw --csv > gpurun_out/trsv_sweep_raw.csv 2>&1
python tools/trace_chol.py 20000 > gpurun_out/trace_chol20k_r2b.txt 2>&1
ls -la gpurun_out | tail -15
