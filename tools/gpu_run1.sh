mkdir -p gpurun_out
python -m pytest tests/test_bc6h_gpu.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_bc6h.log
python tools/time_format.py BC6HU 2>&1 | tail -1 | tee gpurun_out/time_bc6hu.json
python tools/time_format.py BC6HS 2>&1 | tail -1 | tee gpurun_out/time_bc6hs.json
bash tools/prof_one.sh BC6HU bc6h_encode bc6hu_r2b
