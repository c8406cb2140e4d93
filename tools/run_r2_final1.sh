python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"; tail -c 500 gpurun_out/bench_r2_n1.err
( time python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err ) 2>&1 | tail -3; echo "ref rc=$?"; tail -c 500 gpurun_out/bench_r2_ref.err
python tools/bench_configs.py 2>&1 | tail -1
