python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "dgemm or chol" 2>&1 | tail -3
python tools/bench_gemm.py 2>&1 | tail -7
python tools/time_chol.py 20000,10000 2>&1 | tail -1
python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -3
python tools/bench_gemm.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_dmma -c 8 --csv --log-file gpurun_out/traffic_gemm8192_r2.csv python tools/bench_gemm.py > gpurun_out/ncu_gemm.log 2>&1
tail -12 gpurun_out/traffic_gemm8192_r2.csv | cut -c1-260
