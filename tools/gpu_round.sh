#!/bin/bash
# One GPU-box pass: all gpu tests, smoke, bench (both arms), ncu launch list + full captures of the three search kernels and of
# the remaining kernel families.  usage: gpu_round.sh [noprof]
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
python bench.py 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json
for f in BC1 BC3 BC4U BC5U ETC1 ETC2 ETC2_ALPHA BC6HS ETC2_PUNCHTHROUGH; do
  python tools/time_format.py $f 2>&1 | tail -1 | tee -a gpurun_out/time_formats.jsonl
done
python tools/time_tiler.py 2>&1 | tail -2 | tee gpurun_out/time_tiler.jsonl
python tools/time_decode.py 2>&1 | tail -3 | tee gpurun_out/time_decode.jsonl
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
bash tools/prof_one.sh BC7 bc7_encode bc7_r2f
bash tools/prof_one.sh BC6HU bc6h_encode bc6hu_r2f
bash tools/prof_one.sh ETC2_RGBA etc_encode etc2_r2f
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_image|untile_blocks|decode_kernel|s3tc_encode|eac_encode' --launch-skip 0 -f -o gpurun_out/misc_r2f python tools/prof_misc.py > gpurun_out/ncu_misc_r2f.log 2>&1
tail -n 2 gpurun_out/ncu_misc_r2f.log
fi
