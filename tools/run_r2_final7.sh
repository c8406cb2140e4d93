# last check of round 2: full GPU tier on the final tree, ncu --set full of the Gram kernel fused with the equilibration pass
python -m pytest tests/ -m gpu -q 2>&1 | tail -3
export TMPD=/tmp/lgpprof; mkdir -p $TMPD
python tools/one_step.py 20000 2 > gpurun_out/plain_step.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gram_fast3 -s 1 -c 1 -o $TMPD/prof_gram_prep python tools/one_step.py 20000 2 > gpurun_out/ncu_gram_prep.log 2>&1
ncu -i $TMPD/prof_gram_prep.ncu-rep --page details > gpurun_out/gram_fast3_prep_details.txt 2>&1
ncu -i $TMPD/prof_gram_prep.ncu-rep --page raw --csv > gpurun_out/gram_fast3_prep_raw.csv 2>&1
