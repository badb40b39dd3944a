NG=${NG:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/s17_bench_n$NG.json 2> gpurun_out/s17_bench_n$NG.err
echo "rc=$?"; tail -c 600 gpurun_out/s17_bench_n$NG.err
