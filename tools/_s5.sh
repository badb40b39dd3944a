set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s10_pytest.log
timeout 200 python tools/profile_step.py --workload c3 > gpurun_out/s10_profile_c3.txt 2>&1
timeout 300 python tools/profile_step.py --workload c4 --aggregate > gpurun_out/s10_profile_c4.txt 2>&1
timeout 200 python bench.py --workload c3 --steps 20 --warmup 5 --cpu-seconds 0 --extras none > gpurun_out/s10_bench_c3.json 2> gpurun_out/s10_bench_c3.err
timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --cpu-seconds 0 --extras none > gpurun_out/s10_bench_c4.json 2> gpurun_out/s10_bench_c4.err
tail -3 gpurun_out/s10_pytest.log; tail -3 gpurun_out/s10_profile_c3.txt; tail -2 gpurun_out/s10_profile_c4.txt
