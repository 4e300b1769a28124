mkdir -p gpurun_out
python -m pytest tests/test_bc7_gpu.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_bc7.log
for n in 65536 75576; do
python tools/time_format.py BC7 $n 2>&1 | tail -1 | cut -c1-110
CVTTB200_BC7_TAIL=0 python tools/time_format.py BC7 $n 2>&1 | tail -1 | cut -c1-110
done | tee gpurun_out/tail_ab.txt
