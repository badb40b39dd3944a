timeout 300 python -m pytest tests/test_gemm_grouped_gpu.py -m gpu -x -q 2>&1 | tail -3
for c in fwd bwd; do TCR_RNN_DEBUG=1 timeout 100 python tools/rnn_gemm_bench.py $c 2 2>&1 | tail -2; done
timeout 120 python tools/rnn_seq_check.py 128 2 2>&1 | tail -1
run() { timeout 200 python bench.py --workload $1 --steps $2 --warmup 3 --cpu-seconds 0 --extras none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['launches_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"; }
echo c4; run c4 5
