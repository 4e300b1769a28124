cat > /tmp/prof_small.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, numpy as np
from convectionkernels_b200 import api, synth
fmt = sys.argv[1]
api.init(0)
n = 151552 if not fmt.startswith("ETC") else 303104      # ETC: persistent grid of 75776 threads, four tiles per CTA
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
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/$3 python /tmp/prof_small.py $1 > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
