# 8 GPUs: the driver's SCALE command at N = 8, plus the multi-rank DistChol tests that need 4 and 8 GPUs
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; echo "bench rc=$?"; tail -c 800 gpurun_out/bench_r2_n8.err
python -m pytest tests/test_gpu_dist.py -m gpu -q -k "multi_rank and (2x2 or 2x4)" 2>&1 | tail -4
