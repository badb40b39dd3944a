for c in fwd bwd; do TCR_RNN_DEBUG=1 timeout 100 python tools/rnn_gemm_bench.py $c 2 2>&1 | tail -2; done
run() { timeout 200 python bench.py --workload $1 --steps $2 --warmup 3 --cpu-seconds 0 --extras none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['launches_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"; }
echo c4; run c4 5
echo c3; run c3 20
