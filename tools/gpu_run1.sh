mkdir -p gpurun_out
python tools/time_format.py BC7 2>&1 | tail -1 | cut -c1-120
python tools/time_format.py BC7 262144 2>&1 | tail -1 | cut -c1-120
