# sweep of the tail item classes (pairs of half-split tiles x pairs of solo tiles) on one box; see tools/ab_xsplit.py
for xs in 0 18 37 74; do for so in 37 74 111 148; do
BLADE_XSPLIT_PAIRS=$xs BLADE_SOLO_PAIRS=$so python tools/ab_xsplit.py 2>&1 | grep -v "^\[W"
done; done
