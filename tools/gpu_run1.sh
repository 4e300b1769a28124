mkdir -p gpurun_out
: > gpurun_out/threads_small.jsonl
python tools/time_threads_small.py 8 | tee -a gpurun_out/threads_small.jsonl
for s in 48 12 6 3; do CVTTB200_BC7_SPLIT=$s python tools/time_threads_small.py 8 | tee -a gpurun_out/threads_small.jsonl; done
CVTTB200_BC7_SPLIT=0 python tools/time_threads_small.py 8 | tee -a gpurun_out/threads_small.jsonl
