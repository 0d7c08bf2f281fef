python bench.py --steps 10 --warmup 3 > gpurun_out/bench9.log 2>gpurun_out/bench9.err; tail -3 gpurun_out/bench9.err; python -c "
import json;d=json.loads(open('gpurun_out/bench9.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']);print(d['roofline']);print(d.get('cpu_baseline'));
[print(k,{a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()}) for k,v in d['kernels_in_step'].items()]
[print(k,round(v['ms'],4),round(v['achieved'],1),round(v['frac'],3)) for k,v in d['ops_isolated'].items()]"
python tests/conv_probe.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"qconv_wgrad_kernel|qconv_igemm" -s 8 -c 30 -o gpurun_out/prof_conv4 python tests/conv_probe.py > gpurun_out/ncu4.log 2>&1
