mkdir -p gpurun_out
python -m pytest tests/test_bc7_gpu.py tests/test_concurrency_gpu.py tests/test_dropin_cpp.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pytest_bc7.log
python tools/time_format.py BC7 2>&1 | tail -1 | tee gpurun_out/time_bc7.json
bash tools/prof_one.sh BC7 bc7_encode bc7_r2a
