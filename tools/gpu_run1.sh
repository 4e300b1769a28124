mkdir -p gpurun_out
python -m pytest tests/test_concurrency_gpu.py tests/test_bc7_gpu.py -q -m gpu 2>&1 | tail -3
python tools/time_small_calls.py BC7 28424 32768 40960 | cut -c1-300
