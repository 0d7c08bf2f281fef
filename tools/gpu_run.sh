timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
QUAN_TC_EPI_STATS=2 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "epilogue or fused" 2>&1 | tail -3
timeout 300 python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
QUAN_TC_EPI_STATS=0 timeout 300 python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table | tail -1 | cut -c1-200
