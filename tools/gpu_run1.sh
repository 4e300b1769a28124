mkdir -p gpurun_out
python -m pytest tests/test_etc_gpu.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_etc.log
: > gpurun_out/small_calls_etc.jsonl
for f in ETC2_RGBA ETC2 ETC1; do python tools/time_small_calls.py $f 8 64 512 2560 4096 8192 16384 32768 37888 | tee -a gpurun_out/small_calls_etc.jsonl; done
CVTTB200_ETC_SPLIT=0 python tools/time_small_calls.py ETC2_RGBA 8 512 4096 32768 | tee -a gpurun_out/small_calls_etc.jsonl
for u in $(seq 0 26); do echo -n "unit $u " | tee -a gpurun_out/small_calls_etc.jsonl; CVTTB200_ETC_ONLY_UNIT=$u python tools/time_small_calls.py ETC2_RGBA 512 | cut -c1-120 | tee -a gpurun_out/small_calls_etc.jsonl; done
for u in 0 1 2 3; do echo -n "etc1 unit $u " | tee -a gpurun_out/small_calls_etc.jsonl; CVTTB200_ETC_ONLY_UNIT=$u python tools/time_small_calls.py ETC1 512 | cut -c1-120 | tee -a gpurun_out/small_calls_etc.jsonl; done
python tools/time_format.py ETC2_RGBA | tee -a gpurun_out/small_calls_etc.jsonl
python tools/time_format.py ETC1 | tee -a gpurun_out/small_calls_etc.jsonl
