set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 300 python -m pytest tests/test_gemm_shapes_gpu.py tests/test_gemm_tc_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "gemm+train tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
for wl in c4 c2 c5 c1; do
  timeout 300 python bench.py --workload $wl --steps 20 --cpu-seconds 3 > gpurun_out/r1_bench_final_$wl.log 2> gpurun_out/r1_bench_final_$wl.err; tail -1 gpurun_out/r1_bench_final_$wl.log | cut -c1-330; tail -2 gpurun_out/r1_bench_final_$wl.err
  echo "bench $wl done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
done
timeout 200 python bench.py --mode inference --steps 30 --cpu-seconds 3 > gpurun_out/r1_bench_final_c3_inference.log 2> gpurun_out/r1_bench_final_c3_inference.err; tail -1 gpurun_out/r1_bench_final_c3_inference.log | cut -c1-1200; tail -3 gpurun_out/r1_bench_final_c3_inference.err
echo "done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
