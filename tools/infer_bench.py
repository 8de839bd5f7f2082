#!/usr/bin/env python
"""Sliding-window inference throughput (BASELINE.json configs[3]): predict_3D on a synthetic 512^3 volume (or smaller),
patch 192x160x128, step 0.5, Gaussian aggregation; reports tiles/s = 3D patches/s of the inference forward + aggregation,
s/volume, and the per-kernel breakdown (CUDA events around every launch in a second pass)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from multitalent_b200 import _lib as L  # noqa: E402
from multitalent_b200.plans import default_plans  # noqa: E402
from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vol", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--patch", type=int, nargs=3, default=[192, 160, 128])
    ap.add_argument("--tta", action="store_true")
    ap.add_argument("--tb", type=int, default=0, help="tiles per network pass (0 = the network's default)")
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
    patch = tuple(a.patch)
    tr = MultiTalent_trainer_ddp(default_plans(patch_size=patch, batch_size=1), 0, 0, native_dtype=dt,
                                 init_distributed=False)
    torch.manual_seed(0)
    tr.initialize(False)
    net = tr.network
    net.eval()
    net.do_ds = False
    if a.tb:
        net.inference_tile_batch = a.tb
    rng = np.random.RandomState(0)
    vol = rng.randn(1, *a.vol).astype(np.float32)
    steps = net._compute_steps_for_sliding_window(patch, tuple(a.vol), 0.5)
    ntiles = len(steps[0]) * len(steps[1]) * len(steps[2]) * (8 if a.tta else 1)
    kw = dict(do_mirroring=a.tta, mirror_axes=(0, 1, 2), use_sliding_window=True, step_size=0.5, patch_size=patch,
              regions_class_order=tuple(range(47)), use_gaussian=True, verbose=False, return_device_tensors=True)
    small = vol[:, :patch[0], :patch[1], :patch[2] + 64]
    s0, p0 = net.predict_3D(vol, **kw)  # warm-up on the same volume (allocator, Gaussian map)
    del s0, p0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seg, prob = net.predict_3D(vol, **kw)
    torch.cuda.synchronize()
    dtm = time.perf_counter() - t0
    del seg, prob
    torch.cuda.empty_cache()
    with L.KernelProfile() as kp:
        net.predict_3D(small, **kw)
    ks = kp.summary()
    tot = sum(v["ms"] for v in ks.values())
    print(json.dumps({"metric": "sliding-window inference", "volume": a.vol, "patch": list(patch), "tiles": ntiles,
                      "tta": a.tta, "dtype": a.dtype, "seconds_per_volume": dtm, "patches_per_s": ntiles / dtm,
                      "includes": "H2D of the volume, 210 x (gather + forward + Gaussian scatter-add), normalise + threshold",
                      "kernel_share_2_tiles": {k: round(v["ms"] / tot, 3) for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms"])[:8]},
                      "ms_per_tile_kernels": tot / 2}))


if __name__ == "__main__":
    main()
