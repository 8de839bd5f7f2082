import sys, torch, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import build_small_net
import json
blob = np.load("tests/golden/generic_small.npz"); meta = json.load(open("tests/golden/generic_small.json"))
from multitalent_b200.engine import Engine
x = torch.from_numpy(blob["x"]).cuda()
orig = Engine.use_c1
for dt in (torch.float16, torch.bfloat16):
    outs = {}
    for mode in ("c1", "noc1"):
        Engine.use_c1 = orig if mode == "c1" else (lambda self, op: False)
        net = build_small_net(meta, blob, dtype=dt)
        with torch.no_grad():
            o = net(x)
        outs[mode] = [t.float().cpu() for t in o]
        print(dt, mode, [round(float((t - torch.from_numpy(blob["logits_%d" % i])).abs().max()), 5) for i, t in enumerate(outs[mode])])
    print(dt, "c1 vs noc1", [round(float((a - b).abs().max()), 5) for a, b in zip(outs["c1"], outs["noc1"])])
Engine.use_c1 = orig
