mkdir -p gpurun_out
python -m pytest tests/test_bc7_gpu.py tests/test_bc6h_gpu.py -q -m gpu -x -k "every_slice_count" 2>&1 | tail -5 | tee gpurun_out/pytest_slices.log
