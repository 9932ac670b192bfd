for r in 16 8 4; do for h in 12 3; do
BLADE_FUSED_ROWS=$r python tools/layer_stages.py --heads $h --tag rows=$r 2>&1 | grep -v "^\[W"
done; BLADE_FUSED_ROWS=$r python tools/layer_stages.py --heads 48 --workload cog --tag rows=$r 2>&1 | grep -v "^\[W"
BLADE_FUSED_ROWS=$r python tools/layer_stages.py --heads 12 --workload cog --tag rows=$r 2>&1 | grep -v "^\[W"
done
python tools/layer_stages.py --heads 3 --tag default; python tools/layer_stages.py --heads 12 --tag default
