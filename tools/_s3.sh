NG=${NG:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $1 "${@:2}" 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS\|^NCCL version"; }
run 29544 tools/p2p_allreduce_check.py > gpurun_out/s3_p2p.log 2>&1
TCR_P2P_ALLREDUCE=0 run 29545 tools/p2p_allreduce_check.py > gpurun_out/s3_nccl.log 2>&1
for mode in 1 0; do
  for w in c3 c4; do
    TCR_P2P_ALLREDUCE=$mode run 2955$mode bench.py --gpus $NG --workload $w --steps 10 --warmup 3 --cpu-seconds 0 --extras none > gpurun_out/s3_bench_${w}_p2p${mode}.log 2>&1
  done
done
tail -c 600 gpurun_out/s3_bench_c3_p2p1.log
