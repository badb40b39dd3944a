mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_gemm_shapes_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s8_pytest.log
tail -3 gpurun_out/s8_pytest.log
for shape in "8192 1024 784 0 0" "4096 4096 4096 0 0" "8192 1024 1024 0 0"; do
  set -- $shape
  for prec in 2 1; do
    for sk in 1 0; do
    TCR_TC2_STREAMK=$sk TCR_TC2_DEBUG=1 timeout 60 python tools/one_gemm.py --m $1 --n $2 --k $3 --ta $4 --tb $5 --prec $prec --iters 20 --warmup 3 --graph 2>&1 | sort -u | tail -2
    done
  done
done > gpurun_out/s8_gemm.txt 2>&1
cat gpurun_out/s8_gemm.txt
