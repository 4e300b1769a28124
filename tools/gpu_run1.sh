python tools/time_format.py BC7 2>&1 | tail -1 | cut -c1-110
CVTTB200_BC7_PAIR_ORDER=0 python tools/time_format.py BC7 2>&1 | tail -1 | cut -c1-110
