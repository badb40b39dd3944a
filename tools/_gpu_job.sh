set +e
timeout 400 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q 2>&1 | tail -8
for wl in c3 c4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/r1_bench_2gpu_$wl.log 2>gpurun_out/r1_bench_2gpu_$wl.err; echo "rc=$?"; tail -1 gpurun_out/r1_bench_2gpu_$wl.log | cut -c1-330; tail -3 gpurun_out/r1_bench_2gpu_$wl.err | cut -c1-300
done
