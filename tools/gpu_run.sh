timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tests/conv_probe.py | tail -1
QUAN_TC_HALO=0 python tests/conv_probe.py | tail -1
for c in 64 128 512; do C=$c HW=$((8192/c)) python tests/conv_probe.py | tail -1; done
DT=f32 python tests/conv_probe.py | tail -1
