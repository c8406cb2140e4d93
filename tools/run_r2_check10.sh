python tools/dist_check.py --size 40960 --tile 1024 --storage lower --reps 2 2>&1 | tail -2 | cut -c1-700
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"; tail -c 800 gpurun_out/bench_r2c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2c.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['phases_ms'], d['chol_alone_ms'], d['roofline']['frac'], d['phase_rates'])
print(json.dumps(d['dist_chol'])[:1800])
PY
