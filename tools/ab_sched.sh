for w in cog wan; do
for env in "" "BLADE_STATIC_SCHED=1" "BLADE_NO_SPLIT=1" "BLADE_STATIC_SCHED=1 BLADE_NO_SPLIT=1"; do
env $env timeout 200 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w [$env]', round(d['ms_per_step'],4), 'attention', round(d['config']['stage_ms']['attention'],4))"
done; done
