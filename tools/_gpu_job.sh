set +e
timeout 400 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q 2>&1 | tail -25 | cut -c1-400
for wl in c3 c5; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/r1_bench_2gpu_$wl.log 2>gpurun_out/r1_bench_2gpu_$wl.err; echo "rc=$?"; tail -1 gpurun_out/r1_bench_2gpu_$wl.log | cut -c1-250; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r1_bench_2gpu_$wl.err | tail -3 | cut -c1-300
done
