# round 2 profiles: per-kernel ncu captures of the Gram / VJP / BART kernels and the launch list of a short bench
set -x
python tools/microbench_r2.py 20000 > gpurun_out/plain_micro.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_fast3 -s 3 -c 1 -o gpurun_out/prof_r2_gram_fast3 python tools/microbench_r2.py 20000 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_fast_vjp -s 3 -c 1 -o gpurun_out/prof_r2_gram_vjp python tools/microbench_r2.py 20000 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_bart -s 3 -c 1 -o gpurun_out/prof_r2_gram_bart python tools/microbench_r2.py 20000 > gpurun_out/ncu3.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 0 --c3-per-gpu 0 --c1 0 --dist-n1 0 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 0 --c3-per-gpu 0 --c1 0 --dist-n1 0 > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -12
