timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python tools/host_profile.py 2>&1 | head -1
timeout 300 python bench.py --workload yolo11n_trace --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-200
