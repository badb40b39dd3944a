set +e
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --workload c2 --steps 50 --cpu-seconds 1 > gpurun_out/r1_bench12_c2.log 2>gpurun_out/r1_bench12_c2.err; tail -1 gpurun_out/r1_bench12_c2.log | cut -c1-1400; tail -3 gpurun_out/r1_bench12_c2.err
