set +e
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "--m 8192 --n 10 --k 1024" "--m 10 --n 1024 --k 8192 --ta 1" "--m 10 --n 9 --k 4096 --ta 1"; do
  for v in 0 1; do echo "vec=$v"; TCR_SKINNY_VEC=$v timeout 60 python tools/one_gemm.py $cfg --prec 2 --iters 20; done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:skinny_rk2 -c 1 -o gpurun_out/r1_ncu_rk2 -f python tools/one_gemm.py --m 8192 --n 10 --k 1024 --prec 2 --iters 1 --warmup 1 > gpurun_out/r1_ncu_rk2.log 2>&1
timeout 200 python tools/profile_step.py --workload c3 > gpurun_out/r1_step_c3_fuse.txt 2>&1; tail -24 gpurun_out/r1_step_c3_fuse.txt
timeout 200 python tools/profile_step.py --workload c5 --aggregate > gpurun_out/r1_step_c5b.txt 2>&1; tail -40 gpurun_out/r1_step_c5b.txt | head -14
timeout 400 python tools/profile_step.py --workload c4 --aggregate --repeats 2 > gpurun_out/r1_step_c4b.txt 2>&1; tail -50 gpurun_out/r1_step_c4b.txt | head -24
for wl in c3 c1 c5 c2; do
timeout 200 python bench.py --workload $wl --steps 50 --cpu-seconds 1 > gpurun_out/r1_bench5_$wl.log 2>gpurun_out/r1_bench5_$wl.err; tail -1 gpurun_out/r1_bench5_$wl.log | cut -c1-300; tail -3 gpurun_out/r1_bench5_$wl.err
done
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-seconds 1 > gpurun_out/r1_bench5_c4.log 2>gpurun_out/r1_bench5_c4.err; tail -1 gpurun_out/r1_bench5_c4.log | cut -c1-300; tail -3 gpurun_out/r1_bench5_c4.err
