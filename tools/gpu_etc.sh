#!/bin/bash
# ETC round on the GPU box: ETC / tiler / BC6H / S3TC parity tests, then device timings of every ETC colour format.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_etc_gpu.py tests/test_bc6h_gpu.py tests/test_s3tc_gpu.py tests/test_bc7_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_etc.log
for f in ETC2_PUNCHTHROUGH ETC2 ETC2_RGBA ETC1; do
  timeout 600 python tools/time_format.py $f 2>&1 | tail -1 | tee -a gpurun_out/time_etc.jsonl
done
