#!/usr/bin/env python
"""Micro-benchmark of the InstanceNorm/LeakyReLU streaming kernels at the full-resolution layer shape (HBM roofline)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multitalent_b200 import _lib as L  # noqa: E402
from multitalent_b200.engine import Engine, Feat, IN_EPS  # noqa: E402


def main():
    dt = torch.bfloat16
    B, D, H, W, Cc = 4, 192, 160, 128, 32
    if len(sys.argv) > 1:
        D, H, W, Cc = [int(a) for a in sys.argv[1:5]]
    eng = Engine(dt, 0)
    y = Feat(torch.randn(B, D, H, W, Cc, device="cuda").to(dt), 0, Cc, Cc)
    g = torch.randn(B, D, H, W, Cc, device="cuda").to(dt)
    out = Feat(torch.empty_like(y.buf), 0, Cc, Cc)
    stats = torch.zeros(B, Cc, 2, dtype=torch.float64, device="cuda")
    L.call("mtb200_in_stats", y.ptr(), L.dtype_enum(dt), B, y.nvox, y.ldc, 0, Cc, L.ptr(stats), L.stream_ptr())
    gamma = torch.ones(Cc, device="cuda"); beta = torch.zeros(Cc, device="cuda")
    eng.finalize_norm(y, stats, gamma, beta)
    red = torch.zeros(B, Cc, 2, dtype=torch.float64, device="cuda")
    dg = torch.zeros(Cc, device="cuda"); db = torch.zeros(Cc, device="cuda")
    nbytes = y.buf.numel() * 2
    d = L.dtype_enum(dt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def k_norm():
        eng.materialize(y, out=out)

    def k_red():
        L.call("mtb200_in_bwd_reduce", g.data_ptr(), Cc, 0, y.ptr(), y.ldc, 0, d, B, y.nvox, Cc, L.ptr(y.xform),
               L.ptr(y.meanrstd), L.ptr(red), L.stream_ptr())

    def k_apply():
        L.call("mtb200_in_bwd_apply", g.data_ptr(), Cc, 0, y.ptr(), y.ldc, 0, g.data_ptr(), Cc, 0, d, B, y.nvox, Cc,
               L.ptr(y.xform), L.ptr(y.meanrstd), L.ptr(gamma), L.ptr(red), L.ptr(dg), L.ptr(db), L.stream_ptr())

    g2 = torch.empty_like(g)

    def k_apply_oop():
        L.call("mtb200_in_bwd_apply", g.data_ptr(), Cc, 0, y.ptr(), y.ldc, 0, g2.data_ptr(), Cc, 0, d, B, y.nvox, Cc,
               L.ptr(y.xform), L.ptr(y.meanrstd), L.ptr(gamma), L.ptr(red), L.ptr(dg), L.ptr(db), L.stream_ptr())

    for name, fn, passes in (("norm_act", k_norm, 2), ("in_bwd_reduce", k_red, 2), ("in_bwd_apply", k_apply, 3),
                             ("apply_outofplace", k_apply_oop, 3)):
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        print("%-14s NU=%s waves=%s  %.3f ms  %.0f GB/s (algorithmic %d passes of %.2f GB)" % (
            name, os.environ.get("MTB200_NORM_NU", "dflt"), os.environ.get("MTB200_NORM_WAVES", "dflt"), ms,
            passes * nbytes / ms / 1e6, passes, nbytes / 1e9))


if __name__ == "__main__":
    main()
