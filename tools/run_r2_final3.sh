# final single-GPU validation of round 2 (third session): full GPU test tier, smoke, bench line with the driver's command
python -m pytest tests/ -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_n1_final.json 2> gpurun_out/bench_r2_n1_final.err
tail -c 600 gpurun_out/bench_r2_n1_final.err
