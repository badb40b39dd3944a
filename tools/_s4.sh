run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}" 2>&1 | grep "^{\|Error\|error\|fatal" | head -5; }
echo "fused small"; run 29561 tools/dp_check_lstm.py 32 64 10 8
echo "unfused small"; TCR_NO_RNN_FUSE=1 run 29562 tools/dp_check_lstm.py 32 64 10 8
echo "fused full"; run 29563 tools/dp_check_lstm.py 128 1024 128 64
echo "unfused full"; TCR_NO_RNN_FUSE=1 run 29564 tools/dp_check_lstm.py 128 1024 128 64
