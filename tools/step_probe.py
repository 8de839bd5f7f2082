#!/usr/bin/env python
"""Per-step host and device time of the first N training steps (allocator warm-up, loss-scale transients)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multitalent_b200.plans import default_plans  # noqa: E402
from multitalent_b200.synthetic import synthetic_batch  # noqa: E402
from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--steps", type=int, default=12)
a = ap.parse_args()
dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
patch = (192, 160, 128)
tr = MultiTalent_trainer_ddp(default_plans(patch_size=patch, batch_size=4), 0, 0, native_dtype=dt, init_distributed=False)
torch.manual_seed(0)
tr.initialize(True)
b = synthetic_batch(patch, 4, 0, tr.deep_supervision_scales)
x = torch.from_numpy(b['data']).cuda()
tg = [torch.from_numpy(t).cuda() for t in b['target']]
valid = [p['valid_regions'] for p in b['properties']]
for i in range(a.steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    l, _, _ = tr.train_step(x, tg, valid, True)
    e1.record()
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    sc = tr.amp_grad_scaler.state.cpu().tolist() if tr.amp_grad_scaler is not None else None
    print("step %2d host %.1f ms device %.1f ms loss %.4f scaler %s reserved %.1f GB" %
          (i, th * 1e3, e0.elapsed_time(e1), float(l), sc, torch.cuda.memory_reserved() / 2 ** 30), flush=True)
