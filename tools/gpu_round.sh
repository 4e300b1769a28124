#!/bin/bash
# One GPU-box pass: gpu tests, smoke, bench (both arms), ncu launch list + full capture of the BC7 kernel.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
cat > /tmp/prof_small.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, numpy as np
from convectionkernels_b200 import api, synth
api.init(0)
blocks = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))[:151552*2]   # 2 waves of 148 SMs x 4 CTAs x 128 threads... x2
d = torch.from_numpy(blocks).cuda()
o, p = api.Options(), api.BC7EncodingPlan(); api.ConfigureBC7EncodingPlanFromQuality(p, 100)
for _ in range(2):
    api.encode("BC7", d, o, p)
torch.cuda.synchronize()
PY
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 1 -c 1 -f -o gpurun_out/bc7_prof python /tmp/prof_small.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
