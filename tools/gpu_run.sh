for i in 3 8 17 18 19 20 21 22 23 24; do timeout 120 python tests/tc_probe.py $i 2>&1 | grep -v "^\[.*pick_algo\|sample\|mismatches" | tail -5; done
