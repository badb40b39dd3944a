set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
TCR_GEMM_SHORTK=1 timeout 300 python -m pytest tests/test_gemm_tc_gpu.py tests/test_gemm_shapes_gpu.py tests/test_conv_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
for knob in 0 1; do
  for shape in "65536 64 288 2" "65536 64 288 1" "65536 64 576 2" "16384 128 800 2" "64 1024 1152 2" "8192 128 256 2"; do
    set -- $shape
    echo -n "shortk=$knob "
    TCR_GEMM_SHORTK=$knob python tools/one_gemm.py --m $1 --n $2 --k $3 --prec $4 --iters 20 --warmup 2 --graph
  done
done > gpurun_out/r1_sweep_shortk.txt 2>&1; cat gpurun_out/r1_sweep_shortk.txt
echo "sweep done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
TCR_GEMM_SHORTK=1 timeout 200 python bench.py --workload conv --steps 20 --cpu-seconds 0 2>/dev/null | tail -1 | cut -c1-220
TCR_GEMM_SHORTK=1 timeout 300 python bench.py --workload c4 --steps 20 --cpu-seconds 0 2>/dev/null | tail -1 | cut -c1-220
echo "done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
