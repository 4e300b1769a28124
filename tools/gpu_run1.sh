mkdir -p gpurun_out
for k in 1 2 4 8 1000; do echo -n "sync period $k: "; CVTTB200_BC7_SHAPE_SYNC=$k python tools/time_format.py BC7 2>&1 | tail -1 | cut -c1-120; done | tee gpurun_out/shape_sync.txt
