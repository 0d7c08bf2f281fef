timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
QUAN_TRACE_KERNELS=1 timeout 300 python bench.py --workload yolo11n_trace --steps 3 --warmup 3 2>&1 | grep -E "ms/step" | head -14
