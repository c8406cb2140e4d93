# round 2 profiles, second attempt: captures go to /tmp, only text summaries come back (gpurun_out is limited to 64 MiB)
export TMPD=/tmp/lgpprof; mkdir -p $TMPD
python tools/microbench_r2.py 20000 > gpurun_out/plain_micro.log 2>&1 || exit 1
for k in gram_fast3 gram_fast_vjp gram_bart; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $TMPD/prof_$k python tools/microbench_r2.py 20000 > gpurun_out/ncu_$k.log 2>&1
  ncu -i $TMPD/prof_$k.ncu-rep --page details > gpurun_out/${k}_details.txt 2>&1
  ncu -i $TMPD/prof_$k.ncu-rep --page raw --csv > gpurun_out/${k}_raw.csv 2>&1
  ncu -i $TMPD/prof_$k.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${k}_source.csv.gz
done
python tools/one_step.py 20000 > gpurun_out/plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_step_r2.csv python tools/one_step.py 20000 > gpurun_out/ncu_step.log 2>&1
gzip -f gpurun_out/launches_step_r2.csv
du -sh gpurun_out; ls -la gpurun_out
