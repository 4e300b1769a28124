mkdir -p gpurun_out
python -m pytest tests/test_etc_gpu.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_etc.log
for b in 0 6 7; do
echo "nosync=$b"; CVTTB200_ETC_NOSYNC=$b python tools/time_format.py ETC2_RGBA 2>&1 | tail -1 | tee gpurun_out/time_etc_nosync$b.json
done
python tools/time_format.py ETC1 2>&1 | tail -1
python tools/time_format.py ETC2_PUNCHTHROUGH 2>&1 | tail -1
bash tools/prof_one.sh ETC2_RGBA etc_encode etc2_r2b
