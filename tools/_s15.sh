mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s15_pytest.log
tail -3 gpurun_out/s15_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s15_bench.json 2> gpurun_out/s15_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/s15_bench.err
timeout 200 python tools/profile_step.py --workload c3 > gpurun_out/s15_profile_c3.txt 2>&1
timeout 300 python tools/profile_step.py --workload c4 --aggregate > gpurun_out/s15_profile_c4.txt 2>&1
