# round 2, check 1: GPU tests + N=1 bench with a reduced 1-GPU dist point (quick)
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 2 --dist-n1 40960 --dist-parity-n 8192 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_r2a.err
