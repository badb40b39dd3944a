set +e
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/microbench.py --only ew > gpurun_out/r1_mb_ew3.jsonl 2>gpurun_out/r1_mb_ew3.err; grep fused gpurun_out/r1_mb_ew3.jsonl | cut -c1-150; tail -3 gpurun_out/r1_mb_ew3.err
for cfg in "--m 8192 --n 10 --k 1024" "--m 8192 --n 1024 --k 784"; do timeout 60 python tools/one_gemm.py $cfg --prec 2 --iters 20; done
for wl in c3 c1 c5 c2 c1w; do
timeout 200 python bench.py --workload $wl --steps 50 --cpu-seconds 1 > gpurun_out/r1_bench7_$wl.log 2>gpurun_out/r1_bench7_$wl.err; tail -1 gpurun_out/r1_bench7_$wl.log | cut -c1-300; tail -3 gpurun_out/r1_bench7_$wl.err
done
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-seconds 1 > gpurun_out/r1_bench7_c4.log 2>gpurun_out/r1_bench7_c4.err; tail -1 gpurun_out/r1_bench7_c4.log | cut -c1-300; tail -3 gpurun_out/r1_bench7_c4.err
