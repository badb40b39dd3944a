for n in 192 256 320 384 448 512; do
  for prec in 2 1; do
    TCR_TC2_DEBUG=1 timeout 60 python tools/one_gemm.py --m 18944 --n $n --k 3136 --prec $prec --iters 10 --warmup 2 --graph 2>&1 | sort -u | tail -2
  done
done > gpurun_out/s7_nsweep.txt 2>&1
cat gpurun_out/s7_nsweep.txt
