S=2 python tests/conv_probe.py | tail -1
S=2 C=128 HW=64 python tests/conv_probe.py | tail -1
K=1 python tests/conv_probe.py | tail -1
N=16 python tests/conv_probe.py | tail -1
N=256 C=128 HW=32 python tests/conv_probe.py | tail -1
