mkdir -p gpurun_out
python -m pytest tests/test_bc7_gpu.py tests/test_dropin_cpp.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_bc7.log
python tools/time_format.py BC7 2>&1 | tail -1 | tee gpurun_out/time_bc7.json
python tools/time_small_calls.py BC7 8 512 4096 18816 | tee gpurun_out/small_bc7.json
