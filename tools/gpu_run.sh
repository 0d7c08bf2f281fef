timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
python tests/iqbn_probe.py | tail -1
QUAN_IQBN_RBPS=2 python tests/iqbn_probe.py | tail -1
QUAN_IQBN_RU=1 python tests/iqbn_probe.py | tail -1
