#!/bin/bash
# Quick BC7 iteration on the GPU box: BC7 parity tests + a short bench (value, e2e, bit-exactness of the sampled CPU check).
python -m pytest tests/test_bc7_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BC7 Mblocks/s', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'bit_exact', d['cpu_baseline'].get('bit_exact_vs_gpu'), 'clk', d['clocks'])"
