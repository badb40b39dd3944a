set +e
for cfg in "--m 4096 --n 4096 --k 4096" "--m 2048 --n 2048 --k 2048" "--m 8192 --n 8192 --k 4096" "--m 8192 --n 1024 --k 784" "--m 1024 --n 784 --k 8192 --ta 1"; do
  for prec in 1 2; do timeout 60 python tools/one_gemm.py $cfg --prec $prec --iters 5; done
done
echo "--- cluster cap 4"; TCR_TC2_CLUSTERS=4 timeout 60 python tools/one_gemm.py --prec 1 --iters 2
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --workload c5 --steps 30 --cpu-seconds 2 > gpurun_out/r1_bench_c5.log 2>gpurun_out/r1_bench_c5.err; tail -1 gpurun_out/r1_bench_c5.log | cut -c1-1800; tail -3 gpurun_out/r1_bench_c5.err
timeout 200 python bench.py --workload c3 --steps 30 --cpu-seconds 5 > gpurun_out/r1_bench3_c3.log 2>gpurun_out/r1_bench3_c3.err; tail -1 gpurun_out/r1_bench3_c3.log | cut -c1-2500; tail -3 gpurun_out/r1_bench3_c3.err
