mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench_line.json
