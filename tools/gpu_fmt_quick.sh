#!/bin/bash
# Quick per-format iteration on the GPU box: usage gpu_fmt_quick.sh <FORMAT> <pytest file>
python -m pytest $2 -m gpu -x -q 2>&1 | tail -2
python bench.py --format $1 --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 Mblocks/s', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'bit_exact', d['cpu_baseline'].get('bit_exact_vs_gpu'), 'cpu', round(d['cpu_baseline']['value'],4), 'clk', d['clocks'])"
