mkdir -p gpurun_out
convectionkernels_b200/_build/lane_model | tee gpurun_out/lane_model.json
python -m pytest tests/test_tiler_gpu.py -q -m gpu -x 2>&1 | tail -3
python tools/time_tiler.py 2>&1 | tail -2 | tee gpurun_out/time_tiler.jsonl
