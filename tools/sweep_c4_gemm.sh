# per-launch GPU cost (CUDA-graph replay of 50 dependent launches) of the LSTM (C4) products for different
# split-K floors (TCR_GEMM_SPLIT_MIN_KB = minimum k-blocks of 32 per split; 0 = never split)
for kb in 8 4 2 1 0; do
  for prec in 1 2; do
    for shape in "64 1024 1152 0 0" "64 1152 1024 0 1" "1152 1024 64 1 0" "64 4096 1152 0 0"; do
      set -- $shape
      echo -n "min_kb=$kb "
      TCR_GEMM_SPLIT_MIN_KB=$kb python tools/one_gemm.py --m $1 --n $2 --k $3 --ta $4 --tb $5 --prec $prec --iters 50 --warmup 2 --graph
    done
  done
done
