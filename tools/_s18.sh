mkdir -p gpurun_out
# (a) launch list of the bench command (C3, 2 timed steps): per-launch durations, cold and serialised
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --cpu-seconds 0 --extras none > gpurun_out/r2_launches_c3.out 2>&1
echo "launch list rc=$?"; grep -c gemm_tc2 gpurun_out/r2_launches_c3.csv
# (b) full capture of the dominant kernel (3xTF32 layer-0 product) and its TF32 variant
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -c 2 -o gpurun_out/r2_ncu_gemm_tc2_3x -f python tools/one_gemm.py --m 8192 --n 1024 --k 784 --prec 2 --iters 1 --warmup 1 > /dev/null 2>&1
echo "tc2 rc=$?"
# (c) the recurrent gate kernel and the backward cell kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_rnn_kernel -c 2 -o gpurun_out/r2_ncu_gemm_rnn -f python tools/rnn_gemm_bench.py fwd 2 > /dev/null 2>&1
echo "rnn rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"reduce_cols_v4|row_window_kernel|u8_to_f32_kernel" -c 6 -o gpurun_out/r2_ncu_hbm_kernels -f python tools/_s16b.py > /dev/null 2>&1
echo "hbm rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
