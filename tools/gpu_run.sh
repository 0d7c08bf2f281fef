for ms in 3 2; do QUAN_TC_WG_MINSTAGES=$ms WHICH=dw DT=f32 python tests/conv_probe.py | tail -1; done
for ms in 3 2; do QUAN_TC_WG_MINSTAGES=$ms WHICH=dw python tests/conv_probe.py | tail -1; done
for ms in 3 2; do QUAN_TC_WG_MINSTAGES=$ms WHICH=dw C=512 HW=16 python tests/conv_probe.py | tail -1; done
