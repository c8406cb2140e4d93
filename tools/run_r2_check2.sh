# round 2, check 2: N=2 bench (sharded headline, c3_batch, dist_chol n=150k + parity n=30000)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_r2_n2.err
python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -5
