#!/usr/bin/env python
"""Per-launch table of one training step (CUDA events on the launching stream): which layer shapes cost what.
usage: python tools/layer_profile.py [--patch D H W] [--batch B] [--impl K] > gpurun_out/layers.txt"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multitalent_b200 import _lib as L  # noqa: E402
from multitalent_b200.plans import default_plans  # noqa: E402
from multitalent_b200.synthetic import synthetic_batch  # noqa: E402
from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, nargs=3, default=[192, 160, 128])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
    patch = tuple(a.patch)
    tr = MultiTalent_trainer_ddp(default_plans(patch_size=patch, batch_size=a.batch), 0, 0, native_dtype=dt,
                                 init_distributed=False)
    torch.manual_seed(0)
    tr.initialize(True)
    tr.network._engine.impl = a.impl
    tr.network._engine.overlap_wgrad = False  # per-kernel durations: nothing else shares the SMs
    batch = synthetic_batch(patch, a.batch, 0, tr.deep_supervision_scales)
    valid = [p['valid_regions'] for p in batch['properties']]
    d = torch.from_numpy(batch['data']).cuda()
    t = [torch.from_numpy(x).cuda() for x in batch['target']]
    for _ in range(2):
        tr.train_step(d, t, valid, True)
    torch.cuda.synchronize()
    with L.KernelProfile() as kp:
        tr.train_step(d, t, valid, True)
    rows = kp.per_launch_kernels()
    tot = sum(r[2] for r in rows)
    print("# one training step, patch %s bs %d %s: %d launches, %.2f ms of kernel time" % (patch, a.batch, a.dtype, len(rows), tot))
    agg = {}
    for name, info, ms, fl, nb, kern in rows:
        k = (name + ":" + kern, info)
        g = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
        g[0] += 1; g[1] += ms; g[2] += fl; g[3] += nb
    print("%-44s %-60s %5s %9s %6s %9s %8s" % ("entry:kernel", "Cin_p,Cout_p,grid,taps,is,os", "n", "ms", "%", "TFLOP/s", "GB/s"))
    for (name, info), (n, ms, fl, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %-60s %5d %9.3f %6.1f %9s %8s" % (name, str(info) if info else "", n, ms, 100 * ms / tot,
                                                         ("%.1f" % (fl / ms / 1e9)) if fl else "-",
                                                         ("%.0f" % (nb / ms / 1e6)) if nb else "-"))


if __name__ == "__main__":
    main()
