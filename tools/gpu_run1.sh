mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
bash tools/prof_one.sh BC7 bc7_encode bc7_r2m
python tools/summarize_profile.py r02_bc7_final gpurun_out/bc7_r2m.ncu-rep 151552 gpurun_out/launches.csv bc7 > /dev/null
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_ref.json
python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
python -c "
import json
b=json.load(open('gpurun_out/bench.json'))
print(b['value'], b['e2e']['value'], b['roofline']['frac'], b['latency_8block_ms'], b['cpu_baseline']['bit_exact_vs_gpu'], b['latency_ms_by_blocks_per_call'])
for k,v in b['other_configs'].items(): print(k, v['value'], v['e2e']['value'], v['roofline']['frac'], v['latency_8block_ms'], v['cpu_baseline']['bit_exact_vs_gpu'])
print(b['strong']['value'], b['strong']['sharded_equals_single_gpu'])
"
