timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
C=64 HW=128 python tests/conv_probe.py | tail -1
QUAN_TC_DBUF=0 C=64 HW=128 python tests/conv_probe.py | tail -1
python tests/conv_probe.py | tail -1
