set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_gemm_shapes_gpu.py tests/test_train_gpu.py -m gpu -q --maxfail=20 -k "im2col or split_k or prefetch or conv" 2>&1 | tail -40 > gpurun_out/r1_new_tests.log; tail -15 gpurun_out/r1_new_tests.log
echo "new tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 200 python tools/microbench.py --only conv > gpurun_out/r1_microbench_conv.jsonl 2> gpurun_out/r1_microbench_conv.err; cat gpurun_out/r1_microbench_conv.jsonl | cut -c1-200; tail -3 gpurun_out/r1_microbench_conv.err
echo "microbench done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 200 python bench.py --workload conv --steps 20 --cpu-seconds 0 > gpurun_out/r1_bench_conv.log 2> gpurun_out/r1_bench_conv.err; tail -1 gpurun_out/r1_bench_conv.log | cut -c1-1700; tail -3 gpurun_out/r1_bench_conv.err
echo "bench conv done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 200 python bench.py --steps 30 --cpu-seconds 5 > gpurun_out/r1_bench_c3.log 2> gpurun_out/r1_bench_c3.err; tail -1 gpurun_out/r1_bench_c3.log | cut -c1-2000; tail -3 gpurun_out/r1_bench_c3.err
echo "bench c3 done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 120 python tools/profile_step.py --workload conv > gpurun_out/r1_step_profile_conv.txt 2>&1; tail -16 gpurun_out/r1_step_profile_conv.txt
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r1_gpu_tests.log; tail -6 gpurun_out/r1_gpu_tests.log
echo "all tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
