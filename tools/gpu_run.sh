timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
python tests/iqbn_probe.py | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench20.log 2>/dev/null; python -c "
import json,sys;d=json.loads(open('gpurun_out/bench20.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value']);
print(' '.join(f\"{k}={v['ms_per_launch']*1e3:.1f}\" for k,v in d['kernels_in_step'].items()))"
timeout 300 python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
