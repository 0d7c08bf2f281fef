timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q -k "iqbn or block or fused" 2>&1 | tail -2
python tests/iqbn_probe.py | tail -1
DT=f32 N=32 python tests/iqbn_probe.py | tail -1
