set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r1_gpu_tests.log; tail -5 gpurun_out/r1_gpu_tests.log
echo "all tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
python tools/one_gemm.py --m 65536 --n 64 --k 288 --prec 1 --iters 20 --warmup 2 --graph
timeout 200 python bench.py --workload conv --precision tf32 --steps 20 --cpu-seconds 0 2>/dev/null | tail -1 | cut -c1-200
echo "done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
