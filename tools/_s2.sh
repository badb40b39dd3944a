for pdl in 0 1; do
  export TCR_PDL=$pdl
  echo "== PDL=$pdl"
  python tools/rnn_gemm_bench.py fwd 2; python tools/rnn_gemm_bench.py bwd 2
  for w in c4 c1 c5 c3; do
    python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 0 --extras none 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['ms_per_step'], d['launches_per_step'], d['final_loss'], d['e2e']['ms_per_step'])"
  done
done
export TCR_PDL=1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
