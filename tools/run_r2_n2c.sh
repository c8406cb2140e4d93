python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_n2_final.json 2> gpurun_out/bench_r2_n2_final.err
echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2_n2_final.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_ref_n2.json 2> gpurun_out/bench_r2_ref_n2.err
echo "ref rc=$?"
