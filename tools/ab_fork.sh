for rep in 1 2; do
for m in 0 1 2; do
BLADE_FORK_MODE=$m timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode $m', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['config']['stage_ms'].items()})"
done; done
