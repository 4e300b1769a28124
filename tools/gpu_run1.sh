mkdir -p gpurun_out
: > gpurun_out/small_calls.jsonl
python tools/time_small_calls.py | tee -a gpurun_out/small_calls.jsonl
CVTTB200_BC7_CLASSIFY_MAX=0 python tools/time_small_calls.py | tee -a gpurun_out/small_calls.jsonl
CVTTB200_BC7_CLASSIFY_MAX=48 python tools/time_small_calls.py | tee -a gpurun_out/small_calls.jsonl
python -m pytest tests/test_bc7_gpu.py tests/test_concurrency_gpu.py tests/test_dropin_cpp.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_bc7.log
