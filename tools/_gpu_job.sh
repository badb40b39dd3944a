set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 200 python -m pytest tests/test_dbn.py -m gpu -q 2>&1 | tail -25
echo "dbn done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_dbn.py 2>&1 | tail -5 > gpurun_out/r1_gpu_tests.log; tail -5 gpurun_out/r1_gpu_tests.log
echo "all tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
