mkdir -p gpurun_out
python -m pytest tests/test_bc7_gpu.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/pytest_bc7.log
python tools/time_small_calls.py BC7 18816 18824 24000 28416 28424 | tee gpurun_out/small_bc7_s2.json
CVTTB200_BC7_SPLIT=0 python tools/time_small_calls.py BC7 18824 24000 28416 | tee -a gpurun_out/small_bc7_s2.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
