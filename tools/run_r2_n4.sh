# 4-GPU validation (2x2 grid: process rows AND columns) of the final kernels: multi-rank DistChol tests, then the driver's command
python -m pytest tests/test_gpu_dist.py -m gpu -q -k "4-2x2" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_r2_n4_final.json 2> gpurun_out/bench_r2_n4_final.err
echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2_n4_final.err
