timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
QUAN_BWD_CONCURRENT=0 timeout 300 python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --workload yolo11n_trace --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-200
