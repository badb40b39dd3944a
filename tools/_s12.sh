mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s12_pytest.log
tail -4 gpurun_out/s12_pytest.log
timeout 300 python tools/profile_step.py --workload c4 --aggregate 2>&1 | head -9
run() { timeout 200 python bench.py --workload $1 --steps $2 --warmup 3 --cpu-seconds 0 --extras none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['launches_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"; }
echo c4; run c4 5
echo c4gru; run c4gru 5
echo c3; run c3 20
