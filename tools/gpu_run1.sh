mkdir -p gpurun_out
python -m pytest tests/test_bc6h_gpu.py -q -m gpu -x 2>&1 | tail -3
python tools/time_small_calls.py BC6HU 8 64 512 1536 4096 8192 16384 32768 | cut -c1-420 | tee gpurun_out/bc6h_seed.txt
CVTTB200_BC6H_SEED=0 python tools/time_small_calls.py BC6HU 8 64 512 1536 4096 8192 16384 32768 | cut -c1-420 | tee -a gpurun_out/bc6h_seed.txt
python tools/time_small_calls.py BC6HS 512 4096 16384 | cut -c1-300 | tee -a gpurun_out/bc6h_seed.txt
CVTTB200_BC6H_SEED=0 python tools/time_small_calls.py BC6HS 512 4096 16384 | cut -c1-300 | tee -a gpurun_out/bc6h_seed.txt
