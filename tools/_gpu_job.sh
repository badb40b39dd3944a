set +e
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_c3_v2.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 1 > gpurun_out/r1_launches_c3_v2.log 2>&1
tail -2 gpurun_out/r1_launches_c3_v2.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 4 -c 2 -o gpurun_out/r1_ncu_c3_gemm -f python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/r1_ncu_c3_gemm.log 2>&1
tail -2 gpurun_out/r1_ncu_c3_gemm.log | cut -c1-200
timeout 200 python tools/profile_step.py --workload c3 > gpurun_out/r1_step_c3_v3.txt 2>&1; tail -22 gpurun_out/r1_step_c3_v3.txt
