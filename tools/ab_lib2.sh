# usage: ab_lib2.sh <other .so> [...]  -- A/B of library builds on the same box: Wan (gaussian, mixed) and CogVideoX layers
for rep in 1 2; do
for lib in "" "$@"; do
for cfg in "wan gaussian" "wan mixed" "cog gaussian"; do set -- $cfg
BLADE_ASA_LIB=$lib timeout 300 python bench.py --workload $1 --inputs $2 --steps 30 --warmup 5 --no-cpu-baseline --no-clip 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib [$(basename ${lib:-default})] $1 $2', round(d['ms_per_step'],4), 'attention', round(d['config']['stage_ms']['attention'],4))"
done; done; done
