# single-GPU round-end style run: GPU tests, bench line, ncu launch list of the bench, per-kernel DRAM traffic of one
# lgp_chol_factor call (for roofline.traffic).  Outputs under gpurun_out/.
python -m pytest tests/test_gpu_dist.py -m gpu -x -q --tb=short 2>&1 | tail -25
python bench.py --steps 3 --warmup 3 > gpurun_out/bench4.log 2> gpurun_out/bench4.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench4.err
python -c "
import json;d=json.loads(open('gpurun_out/bench4.log').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['phases_ms'],d.get('batch_throughput'))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_bench_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 0 > gpurun_out/ncu_bench2.log 2>&1; echo "ncu launches rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'gemm_dmma|potrf_leaf|chol_prepare|chol_jitter|copy_block|sym_scale' -c 4000 --csv --log-file gpurun_out/traffic_chol20k.csv python tools/prof_chol.py 20000 factor > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
