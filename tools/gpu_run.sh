timeout 300 python bench.py --dtype f32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f32.log 2>gpurun_out/bench_f32.err; tail -2 gpurun_out/bench_f32.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/bench_f32.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],round(d['roofline']['frac'],3));
print(' '.join(f\"{k}={v['ms_per_launch']*1e3:.1f}\" for k,v in d['kernels_in_step'].items()))"
WHICH=dw python tests/conv_probe.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:qconv_wgrad_kernel" -s 4 -c 1 -o gpurun_out/prof_wgrad2 env WHICH=dw python tests/conv_probe.py > gpurun_out/ncu8.log 2>&1
