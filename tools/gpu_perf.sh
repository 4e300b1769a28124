#!/bin/bash
# Quick perf iteration on the GPU box: golden parity + reference crop parity, bench (our arm), light ncu of the BC7 kernel.
mkdir -p gpurun_out
python -m pytest tests/test_bc7_gpu.py -x -q -m gpu -k "golden or reference_on_this_host or ragged or division or oracle" 2>&1 | tail -5 | tee gpurun_out/pytest_quick.log
python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_quick.json
cat > /tmp/prof_small.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, numpy as np
from convectionkernels_b200 import api, synth
api.init(0)
blocks = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))[:151552*2]
d = torch.from_numpy(blocks).cuda()
o, p = api.Options(), api.BC7EncodingPlan(); api.ConfigureBC7EncodingPlanFromQuality(p, 100)
for _ in range(2):
    api.encode("BC7", d, o, p)
torch.cuda.synchronize()
PY
timeout 600 ncu --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section InstructionStats --clock-control none -k regex:bc7_encode -s 1 -c 1 --csv --page raw --log-file gpurun_out/ncu_light.csv python /tmp/prof_small.py > gpurun_out/ncu_light.log 2>&1
tail -2 gpurun_out/ncu_light.log
