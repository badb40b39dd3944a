set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s1_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/s1_bench_c3.json 2> gpurun_out/s1_bench_c3.err; echo "rc=$?" >> gpurun_out/s1_bench_c3.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/s1_bench_ref.json 2> gpurun_out/s1_bench_ref.err; echo "rc=$?" >> gpurun_out/s1_bench_ref.err
tail -3 gpurun_out/s1_pytest.log; cat gpurun_out/s1_smoke.log | tail -3; cat gpurun_out/s1_bench_c3.err | tail -5
