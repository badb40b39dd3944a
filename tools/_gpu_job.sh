set +e
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_train_gpu.py tests/test_equation_golden.py tests/test_gemm_tc_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/microbench.py --only ew > gpurun_out/r1_mb_ew7.jsonl 2>gpurun_out/r1_mb_ew7.err; grep "fused\|SIGMOID" gpurun_out/r1_mb_ew7.jsonl | cut -c1-150; tail -3 gpurun_out/r1_mb_ew7.err
for wl in c3 c1 c5; do
timeout 200 python bench.py --workload $wl --steps 50 --cpu-seconds 1 > gpurun_out/r1_bench11_$wl.log 2>gpurun_out/r1_bench11_$wl.err; tail -1 gpurun_out/r1_bench11_$wl.log | cut -c1-220; tail -3 gpurun_out/r1_bench11_$wl.err
done
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-seconds 1 > gpurun_out/r1_bench11_c4.log 2>gpurun_out/r1_bench11_c4.err; tail -1 gpurun_out/r1_bench11_c4.log | cut -c1-220; tail -3 gpurun_out/r1_bench11_c4.err
