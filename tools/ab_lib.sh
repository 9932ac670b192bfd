# usage: ab_lib.sh <other .so> [...]   -- A/B of library builds on the same box (attention stage of the Wan bench)
for rep in 1 2; do
for lib in "" "$@"; do
BLADE_ASA_LIB=$lib timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib [${lib:-default}]', round(d['ms_per_step'],4), 'attention', round(d['config']['stage_ms']['attention'],4))"
done; done
