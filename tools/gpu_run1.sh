mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
