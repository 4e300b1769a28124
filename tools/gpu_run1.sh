mkdir -p gpurun_out
python -m pytest tests/test_bc7_gpu.py tests/test_dropin_cpp.py -q -m gpu -x 2>&1 | tail -2
python tools/time_small_calls.py BC7 8 512 1152 4096 9216 18816 28416 65536 | cut -c1-500 | tee gpurun_out/small_bc7_triple.json
