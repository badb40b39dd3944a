set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
TCR_GEMM_SPLIT_MIN_KB=8 timeout 300 python bench.py --workload c4 --steps 20 --cpu-seconds 0 > gpurun_out/r1_bench_c4_kb8.log 2> gpurun_out/r1_bench_c4_kb8.err; tail -1 gpurun_out/r1_bench_c4_kb8.log | cut -c1-230; tail -2 gpurun_out/r1_bench_c4_kb8.err
echo "kb8 done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
TCR_GEMM_SPLIT_MIN_KB=4 timeout 300 python bench.py --workload c4 --steps 20 --cpu-seconds 0 > gpurun_out/r1_bench_c4_kb4.log 2> gpurun_out/r1_bench_c4_kb4.err; tail -1 gpurun_out/r1_bench_c4_kb4.log | cut -c1-230
echo "kb4 done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
TCR_GEMM_SPLIT_MIN_KB=8 TCR_GRAPH_LANES=1 timeout 300 python bench.py --workload c4 --steps 20 --cpu-seconds 0 > gpurun_out/r1_bench_c4_kb8_lane1.log 2>/dev/null; tail -1 gpurun_out/r1_bench_c4_kb8_lane1.log | cut -c1-230
echo "lane1 done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
