python tests/iqbn_probe.py | tail -1
QUAN_IQBN_TMA_BPS=2 python tests/iqbn_probe.py | tail -1
QUAN_IQBN_TMA_BPS=4 QUAN_IQBN_TMA_KB=48 python tests/iqbn_probe.py | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q -k "iqbn" 2>&1 | tail -2
