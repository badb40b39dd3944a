set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r1_gpu_tests.log; tail -6 gpurun_out/r1_gpu_tests.log
echo "all tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 400 bash tools/sweep_c4_gemm.sh > gpurun_out/r1_sweep_c4_gemm.txt 2>&1; cat gpurun_out/r1_sweep_c4_gemm.txt
echo "sweep done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 200 python tools/microbench.py --only conv > gpurun_out/r1_microbench_conv.jsonl 2> gpurun_out/r1_microbench_conv.err
timeout 120 python tools/profile_step.py --workload conv > gpurun_out/r1_step_profile_conv.txt 2>&1
echo "done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
