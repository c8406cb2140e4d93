# 2 GPUs: multi-rank DistChol tests (both storages, peer on/off, operator API), NVLink byte counters around one fused
# (peer-memory) and one NCCL factorisation at n = 61440, then the N = 2 bench
python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -4
for mode in on off; do
  nvidia-smi nvlink -gt d > gpurun_out/nvlink_before_$mode.txt 2>&1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py --size 61440 --tile 1024 --peer $mode 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -2 | cut -c1-600
  nvidia-smi nvlink -gt d > gpurun_out/nvlink_after_$mode.txt 2>&1
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --c3-per-gpu 4 > gpurun_out/bench_r2_n2b.json 2> gpurun_out/bench_r2_n2b.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2_n2b.err
