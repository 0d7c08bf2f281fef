timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "iqbn_streaming" 2>&1 | tail -12
