set +e
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 300 python -m pytest tests/test_dp_nccl_gpu.py tests/test_conv_gpu.py tests/test_dbg_gpu.py -m gpu -q -k "dp or col2im or plugable" 2>&1 | tail -8
echo "tests done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --cpu-seconds 3 > gpurun_out/r1_bench_2gpu_c3.log 2> gpurun_out/r1_bench_2gpu_c3.err; tail -1 gpurun_out/r1_bench_2gpu_c3.log | cut -c1-1400; tail -3 gpurun_out/r1_bench_2gpu_c3.err
echo "bench 2gpu done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 --cpu-seconds 3 2>&1 | tail -2 | cut -c1-600
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload conv --steps 20 --cpu-seconds 0 > gpurun_out/r1_bench_2gpu_conv.log 2> gpurun_out/r1_bench_2gpu_conv.err; tail -1 gpurun_out/r1_bench_2gpu_conv.log | cut -c1-900; tail -3 gpurun_out/r1_bench_2gpu_conv.err
timeout 100 python tools/microbench.py --only conv 2>/dev/null | grep col2im | cut -c1-200
echo "all done $(( $(date +%s) - $(cat gpurun_out/t0) ))s"
