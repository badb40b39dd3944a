timeout 120 python tools/rnn_seq_check.py 8 2 2>&1 | tail -5
timeout 120 python tools/rnn_seq_check.py 128 2 2>&1 | tail -4
timeout 120 python tools/rnn_seq_check.py 128 1 2>&1 | tail -2
