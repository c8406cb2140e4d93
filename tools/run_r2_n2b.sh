# 2-GPU validation after the leaf / GEMM-tile / TRSV changes: multi-rank DistChol tests, then the driver's N = 2 command
python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_n2_final.json 2> gpurun_out/bench_r2_n2_final.err
echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r2_n2_final.err
