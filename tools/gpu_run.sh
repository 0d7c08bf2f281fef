python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tests/conv_probe.py | tail -1
for c in 64 128 512; do C=$c HW=$((8192/c)) python tests/conv_probe.py | tail -1; done
python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
