#!/bin/bash
# One GPU-box pass: all gpu tests, smoke, bench (both arms, headline + the other single-GPU configs), ncu launch list + full
# captures of the three big kernels.  usage: gpu_round.sh [noprof]
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json
for f in BC6HU ETC2_RGBA; do
  python bench.py --format $f --steps 3 2>&1 | tail -1 | tee gpurun_out/bench_$f.json
done
for f in BC1 BC3 BC4U BC5U ETC1 ETC2 ETC2_ALPHA BC6HS; do
  python tools/time_format.py $f 2>&1 | tail -1 | tee -a gpurun_out/time_formats.jsonl
done
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
cat > /tmp/prof_small.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, numpy as np
from convectionkernels_b200 import api, synth
fmt = sys.argv[1]
api.init(0)
n = 151552 * 2
if fmt.startswith("BC6H"):
    blocks = synth.image_to_blocks(synth.hdr_ramp_f16(4096, 4096))[:n]
else:
    blocks = synth.image_to_blocks(synth.mixed_rgba8(4096, 4096))[:n]
d = torch.from_numpy(blocks).cuda()
o, p = api.Options(), None
if fmt == "BC7":
    p = api.BC7EncodingPlan(); api.ConfigureBC7EncodingPlanFromQuality(p, 100)
for _ in range(2):
    api.encode(fmt, d, o, p)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 1 -c 1 -f -o gpurun_out/bc7_prof python /tmp/prof_small.py BC7 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc6h_encode -s 1 -c 1 -f -o gpurun_out/bc6h_prof python /tmp/prof_small.py BC6HU > gpurun_out/ncu_full_bc6h.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:etc_encode -s 1 -c 1 -f -o gpurun_out/etc2_prof python /tmp/prof_small.py ETC2_RGBA > gpurun_out/ncu_full_etc2.log 2>&1
for f in gpurun_out/ncu_full.log gpurun_out/ncu_full_bc6h.log gpurun_out/ncu_full_etc2.log; do tail -n 2 $f; done
fi
