set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_hone.py tests/test_dbg_gpu.py tests/test_interop_gpu.py -m gpu -q --maxfail=20 2>&1 | tail -40 > gpurun_out/r1_new_tests.log; tail -25 gpurun_out/r1_new_tests.log
echo "new tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 200 python tools/microbench.py --only conv > gpurun_out/r1_microbench_conv.jsonl 2> gpurun_out/r1_microbench_conv.err; grep -E "col2im" gpurun_out/r1_microbench_conv.jsonl | cut -c1-200; tail -3 gpurun_out/r1_microbench_conv.err
echo "microbench done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"col2im" -c 1 -o gpurun_out/r1_ncu_col2im python tools/microbench.py --only conv > gpurun_out/r1_ncu_col2im.log 2>&1
python tools/ncu_summary.py gpurun_out/r1_ncu_col2im.ncu-rep > gpurun_out/r1_ncu_col2im_summary.txt 2>&1; head -30 gpurun_out/r1_ncu_col2im_summary.txt
echo "ncu done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r1_gpu_tests.log; tail -6 gpurun_out/r1_gpu_tests.log
echo "all tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
