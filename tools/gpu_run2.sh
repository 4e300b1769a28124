mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>gpurun_out/bench_ref_n2.err | tail -1 | tee gpurun_out/bench_ref_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2.err | tail -1 | tee gpurun_out/bench_n2.json
tail -3 gpurun_out/bench_n2.err
