timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tests/iqbn_probe.py | tail -1
QUAN_IQBN_TMA=0 python tests/iqbn_probe.py | tail -1
C=16 HW=128 N=16 python tests/iqbn_probe.py | tail -1
QUAN_IQBN_TMA=0 C=16 HW=128 N=16 python tests/iqbn_probe.py | tail -1
