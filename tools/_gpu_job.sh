set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r1_gpu_tests.log; tail -6 gpurun_out/r1_gpu_tests.log
echo "all tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for wl in c3 c1 c2 c5 conv; do
  extra="--cpu-seconds 3"; [ $wl = conv ] && extra="--cpu-seconds 0"; [ $wl = c3 ] && extra="--cpu-seconds 10"
  timeout 300 python bench.py --workload $wl --steps 30 $extra > gpurun_out/r1_bench_final_$wl.log 2> gpurun_out/r1_bench_final_$wl.err; tail -1 gpurun_out/r1_bench_final_$wl.log | cut -c1-200; tail -2 gpurun_out/r1_bench_final_$wl.err
  echo "bench $wl done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 --cpu-seconds 3 2>&1 | tail -1 | cut -c1-300
echo "done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
