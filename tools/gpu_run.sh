for i in 5 11 12 21 25 26 27; do timeout 120 python tests/tc_probe.py $i 2>&1 | grep -v "sample\|mismatches" | tail -5; done
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
python bench.py --workload yolo11n_trace --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-200
