timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_final.log 2>gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'],d['gpu_launches'],d['clocks']);print(d['roofline']);print(d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
