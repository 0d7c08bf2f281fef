python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench13.log 2>/dev/null; python -c "
import json,sys;d=json.loads(open('gpurun_out/bench13.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']);
print(' '.join(f\"{k}={v['ms_per_launch']*1e3:.1f}\" for k,v in d['kernels_in_step'].items()))"
python bench.py --workload yolo11n_trace --steps 5 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-200
python __graft_entry__.py smoke 2>&1 | tail -4
