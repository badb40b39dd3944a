set +e
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_gemm_shapes_gpu.py tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "--- gate GEMM in a graph (GPU-side cost): fused reduce on/off, split min kb 8/4"
for fr in 1 0; do for mk in 8 4; do echo "fused=$fr minkb=$mk"; TCR_GEMM_FUSED_REDUCE=$fr TCR_GEMM_SPLIT_MIN_KB=$mk timeout 60 python tools/one_gemm.py --m 64 --n 1024 --k 1152 --prec 2 --iters 200 --graph; done; done
echo "--- no split"; TCR_GEMM_SPLIT_MIN_KB=0 timeout 60 python tools/one_gemm.py --m 64 --n 1024 --k 1152 --prec 2 --iters 200 --graph
echo "--- dW 1152x1024x64"; timeout 60 python tools/one_gemm.py --m 1152 --n 1024 --k 64 --prec 2 --ta 1 --iters 200 --graph
echo "--- c4"
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-seconds 1 2>&1 | tail -1 | cut -c1-220
TCR_GEMM_FUSED_REDUCE=0 timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-seconds 1 2>&1 | tail -1 | cut -c1-220
