python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_factor or device_resident" 2>&1 | tail -15
python -m pytest tests/test_gpu_api.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -5
python bench.py --steps 4 --warmup 2 --dist-n1 0 --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['phases_ms'], d['chol_alone_ms'], d['chol_inverse_span_ms'], d['batch_throughput'], d['c3_batch']['evals_per_s_by_in_flight'], d['c1_latency'])
PY
