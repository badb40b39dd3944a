timeout 120 python tools/rnn_seq_check.py 12 2 48 32 256 2>&1 | tail -3
timeout 120 python tools/rnn_seq_check.py 12 1 48 32 256 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gemm_grouped_gpu.py -m gpu -x -q 2>&1 | tail -3; echo
