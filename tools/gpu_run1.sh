python tools/time_format.py BC7 2>&1 | tail -1 | cut -c1-200
