timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
python tests/conv_probe.py | tail -1
QUAN_TC_WG_HALO=0 WHICH=dw python tests/conv_probe.py | tail -1
for c in 64 128 512; do WHICH=dw C=$c HW=$((8192/c)) python tests/conv_probe.py | tail -1; done
WHICH=dw DT=f32 python tests/conv_probe.py | tail -1
