TCR_TC2_TRACE=1 timeout 120 python tools/one_gemm.py --m 8192 --n 1024 --k 784 --prec 2 --iters 1 --warmup 1 2>&1 | tail -78 > gpurun_out/s9_trace.txt
for shape in "8192 1024 784 0 0" "4096 4096 4096 0 0"; do
  set -- $shape
  for prec in 2 1; do
    for sk in 1 0; do
    TCR_TC2_STREAMK=$sk TCR_TC2_DEBUG=1 timeout 60 python tools/one_gemm.py --m $1 --n $2 --k $3 --ta $4 --tb $5 --prec $prec --iters 20 --warmup 3 --graph 2>&1 | sort -u | tail -2
    done
  done
done
