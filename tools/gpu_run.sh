set -x
for i in 0 7 14 15 16; do timeout 120 python tests/tc_probe.py $i 2>&1 | tail -5; done
timeout 300 python tests/conv_probe.py | tail -1
