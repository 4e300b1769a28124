mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bc6h -c 24 --csv --log-file gpurun_out/bc6h_split_launches.csv python tools/time_small_calls.py BC6HU 8 4096 16384 > /dev/null 2>&1
cut -d, -f5,9,12- gpurun_out/bc6h_split_launches.csv | tail -26
CVTTB200_BC6H_SPLIT=196 python tools/time_small_calls.py BC6HU 4096 16384 | cut -c1-300
CVTTB200_BC6H_SPLIT=98 python tools/time_small_calls.py BC6HU 4096 16384 | cut -c1-300
