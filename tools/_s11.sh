run() { timeout 200 python bench.py --workload c4 --steps 5 --warmup 3 --cpu-seconds 0 --extras none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['launches_per_step'], d['e2e']['ms_per_step'])"; }
echo default; run
echo no_ew_merge; TCR_NO_EW_MERGE=1 run
echo lanes1; TCR_GRAPH_LANES=1 run
echo lanes1_nomerge; TCR_GRAPH_LANES=1 TCR_NO_EW_MERGE=1 run
echo lanes2; TCR_GRAPH_LANES=2 run
echo pdl; TCR_PDL=1 run
echo c3; timeout 200 python bench.py --workload c3 --steps 20 --warmup 5 --cpu-seconds 0 --extras none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['launches_per_step'], d['e2e']['ms_per_step'])"
timeout 100 python tools/profile_step.py --workload c3 2>&1 | head -5
