mkdir -p gpurun_out
bash tools/prof_one.sh ETC2_RGBA etc_encode etc2_r2j
python tools/summarize_profile.py r02_etc2rgba_final gpurun_out/etc2_r2j.ncu-rep 303104 - etc2_rgba > /dev/null
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_ref.json
python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
python -c "
import json
b=json.load(open('gpurun_out/bench.json'))
print(b['value'], b['e2e']['value'], b['roofline']['frac'], b['latency_8block_ms'])
for k,v in b['other_configs'].items(): print(k, v['value'], v['e2e']['value'], v['roofline']['frac'], v['latency_8block_ms'], v['cpu_baseline']['bit_exact_vs_gpu'])
"
