mkdir -p gpurun_out
for f in ETC2_RGBA ETC1; do python tools/time_format.py $f 2>&1 | tail -1 | cut -c1-130; done | tee gpurun_out/etc_giveup.txt
