#!/bin/bash
# One GPU-box pass: all gpu tests, smoke, ncu launch list + full captures of the three search kernels (summarised on the box so
# that the bench line that follows carries the instruction counts of exactly this build), bench (both arms), the other formats,
# captures of the remaining kernel families.  usage: gpu_round.sh [noprof]
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
bash tools/prof_one.sh BC7 bc7_encode bc7_r2h
bash tools/prof_one.sh BC6HU bc6h_encode bc6hu_r2h
bash tools/prof_one.sh ETC2_RGBA etc_encode etc2_r2h
python tools/summarize_profile.py r02_bc7_final gpurun_out/bc7_r2h.ncu-rep 151552 gpurun_out/launches.csv bc7 > /dev/null
python tools/summarize_profile.py r02_bc6hu_final gpurun_out/bc6hu_r2h.ncu-rep 151552 - bc6hu > /dev/null
python tools/summarize_profile.py r02_etc2rgba_final gpurun_out/etc2_r2h.ncu-rep 303104 - etc2_rgba > /dev/null
mkdir -p gpurun_out/profiles_box && cp profiles/*_kernel_ncu_summary.json profiles/r02_*_final_* gpurun_out/profiles_box/
fi
python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
python bench.py 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json
rm -f gpurun_out/time_formats.jsonl
for f in BC1 BC3 BC4U BC5U ETC1 ETC2 ETC2_ALPHA BC6HS ETC2_PUNCHTHROUGH; do
  python tools/time_format.py $f 2>&1 | tail -1 | tee -a gpurun_out/time_formats.jsonl
done
python tools/time_tiler.py 2>&1 | tail -2 | tee gpurun_out/time_tiler.jsonl
python tools/time_decode.py 2>&1 | tail -3 | tee gpurun_out/time_decode.jsonl
if [ "$1" != "noprof" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_image|untile_blocks|decode_kernel|s3tc_encode|eac_encode' --launch-skip 0 -f -o gpurun_out/misc_r2h python tools/prof_misc.py > gpurun_out/ncu_misc_r2h.log 2>&1
tail -n 2 gpurun_out/ncu_misc_r2h.log
fi
