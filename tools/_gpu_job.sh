set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 1 -o gpurun_out/r1_ncu_conv_gemm python tools/one_gemm.py --m 65536 --n 64 --k 288 --prec 2 --iters 2 --warmup 1 > gpurun_out/r1_ncu_conv_gemm.log 2>&1
python tools/ncu_summary.py gpurun_out/r1_ncu_conv_gemm.ncu-rep > gpurun_out/r1_ncu_conv_gemm_summary.txt 2>&1; cat gpurun_out/r1_ncu_conv_gemm_summary.txt | head -30
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 1 -o gpurun_out/r1_ncu_conv_gemm_tf32 python tools/one_gemm.py --m 65536 --n 64 --k 288 --prec 1 --iters 2 --warmup 1 > gpurun_out/r1_ncu_conv_gemm_tf32.log 2>&1
python tools/ncu_summary.py gpurun_out/r1_ncu_conv_gemm_tf32.ncu-rep > gpurun_out/r1_ncu_conv_gemm_tf32_summary.txt 2>&1; cat gpurun_out/r1_ncu_conv_gemm_tf32_summary.txt | head -30
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"col2im" -c 1 -o gpurun_out/r1_ncu_col2im_vec python tools/microbench.py --only conv > gpurun_out/r1_ncu_col2im_vec.log 2>&1
python tools/ncu_summary.py gpurun_out/r1_ncu_col2im_vec.ncu-rep > gpurun_out/r1_ncu_col2im_vec_summary.txt 2>&1; cat gpurun_out/r1_ncu_col2im_vec_summary.txt | head -30
python tools/one_gemm.py --m 65536 --n 64 --k 288 --prec 2 --iters 20 --warmup 2 --graph
python tools/one_gemm.py --m 65536 --n 64 --k 288 --prec 1 --iters 20 --warmup 2 --graph
python tools/one_gemm.py --m 65536 --n 128 --k 288 --prec 2 --iters 20 --warmup 2 --graph
python tools/one_gemm.py --m 65536 --n 256 --k 288 --prec 2 --iters 20 --warmup 2 --graph
python tools/one_gemm.py --m 32768 --n 128 --k 576 --prec 2 --iters 20 --warmup 2 --graph
echo "done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
