python tools/time_format.py BC7 2>&1 | tail -1 | cut -c1-200
python -m pytest tests/test_bc7_gpu.py -q -m gpu -x 2>&1 | tail -2
