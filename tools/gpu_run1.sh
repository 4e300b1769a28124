mkdir -p gpurun_out
for f in ETC2_RGBA ETC2_ALPHA; do python tools/time_format.py $f 2>&1 | tail -1 | cut -c1-200; done | tee gpurun_out/eac_lut.txt
python -m pytest tests/test_etc_gpu.py tests/test_ktx.py tests/test_dropin_cpp.py -q -m gpu -x 2>&1 | tail -3
