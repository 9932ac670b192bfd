# layer-level check of the tail item classes on both workloads (bench.py stage split)
for cfg in "0 74" "18 74" "18 111" "18 148" "37 148"; do set -- $cfg
for w in wan cog; do for inp in gaussian mixed; do
BLADE_XSPLIT_PAIRS=$1 BLADE_SOLO_PAIRS=$2 timeout 300 python bench.py --workload $w --inputs $inp --steps 30 --warmup 5 --no-cpu-baseline --no-clip 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('xs=$1 solo=$2 $w $inp', round(d['ms_per_step'],4), 'attention', round(d['config']['stage_ms']['attention'],4))"
done; done; done
