mkdir -p gpurun_out
python tools/time_format.py BC6HU | tee gpurun_out/time_bc6hu.json
python tools/time_format.py BC6HS | tee gpurun_out/time_bc6hs.json
python -m pytest tests/test_bc6h_gpu.py -q -m gpu -x 2>&1 | tail -3
