python -m pytest tests/ -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --dist-n1 0 --dist-parity-n 0 --no-cpu-baseline > gpurun_out/bench_r2_mid2.json 2> gpurun_out/bench_r2_mid2.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_mid2.err
