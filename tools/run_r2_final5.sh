# final state of round 2: bench line with the driver's command, launch list of one step (ncu), traffic of one factorisation
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_n1_final.json 2> gpurun_out/bench_r2_n1_final.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_n1_final.err
export TMPD=/tmp/lgpprof; mkdir -p $TMPD
python tools/one_step.py 20000 > gpurun_out/plain_step.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_step_r2c.csv python tools/one_step.py 20000 > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_step_r2c.csv > gpurun_out/launches_step_r2c_summary.txt 2>&1
gzip -f gpurun_out/launches_step_r2c.csv
python tools/prof_chol.py 20000 factor > gpurun_out/plain_chol.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $TMPD/traffic.csv python tools/prof_chol.py 20000 factor > gpurun_out/ncu_traffic.log 2>&1
python tools/summarize_traffic.py $TMPD/traffic.csv 'gram_' > gpurun_out/traffic_chol20k_r2c.txt 2>&1
