#!/usr/bin/env python
"""Do a tensor-core kernel and an HBM-bound streaming pass share the SMs productively?  Times, at the full-resolution
layer shape, (a) one convolution kernel alone, (b) one InstanceNorm-backward apply pass alone, (c) both launched on two
streams at the same time (either order).  (c) < (a) + (b) means the overlap of the weight-gradient stream with the
norm passes of the backward chain can pay; (c) == (a) + (b) means they are bound by the same resource.
usage: python tools/overlap_probe.py [wgrad|dgrad|fwd] [layer]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch import nn  # noqa: E402
from multitalent_b200 import _lib as L  # noqa: E402
from multitalent_b200.engine import ConvOp, Engine, Feat, Tape  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "wgrad"
    cin, cout, dims = {"l0b": (30, 30, (4, 192, 160, 128)), "l1b": (60, 60, (4, 96, 80, 64)),
                       "l2b": (120, 120, (4, 48, 40, 32))}[sys.argv[2] if len(sys.argv) > 2 else "l0b"]
    dt = torch.bfloat16
    B, D, H, W = dims
    eng = Engine(dt, 0)
    eng.overlap_wgrad = False
    conv = nn.Conv3d(cin, cout, 3, 1, 1, bias=True).cuda()
    op = ConvOp(conv.weight, conv.bias, (3, 3, 3), (1, 1, 1))
    x = Feat(torch.randn(B, D, H, W, op.Cin_p, device="cuda").to(dt), 0, cin, op.Cin_p)
    dy = Feat(torch.randn(B, D, H, W, op.Cout_p, device="cuda").to(dt), 0, cout, op.Cout_p)
    out = Feat(torch.empty(B, D, H, W, op.Cout_p, device="cuda", dtype=dt), 0, cout, op.Cout_p)
    # operands of the streaming pass (separate tensors of the same size)
    Cc = op.Cout_p
    y = Feat(torch.randn(B, D, H, W, Cc, device="cuda").to(dt), 0, Cc, Cc)
    g = torch.randn(B, D, H, W, Cc, device="cuda").to(dt)
    stats = torch.zeros(B, Cc, 2, dtype=torch.float64, device="cuda")
    L.call("mtb200_in_stats", y.ptr(), L.dtype_enum(dt), B, y.nvox, y.ldc, 0, Cc, L.ptr(stats), L.stream_ptr())
    gamma = torch.ones(Cc, device="cuda"); beta = torch.zeros(Cc, device="cuda")
    eng.finalize_norm(y, stats, gamma, beta)
    red = torch.zeros(B, Cc, 2, dtype=torch.float64, device="cuda")
    dg = torch.zeros(Cc, device="cuda"); db = torch.zeros(Cc, device="cuda")
    d = L.dtype_enum(dt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def k_conv():
        if what == "fwd":
            eng.conv(op, x, out=out, want_stats=True)
        elif what == "dgrad":
            tape = Tape()
            tape.grad_bufs[id(dy.buf)] = dy.buf
            tape.grad_init[id(dy.buf)] = set()
            gx, _ = tape.grad_feat(x)
            eng._conv_call(op.dgrad_taps, dy, op.packed(eng.wdtype, True), None, gx, gx.dims[1:], None, False,
                           op.Cout_p, op.Cin_p)
        else:
            tape = Tape()
            tape.grad_bufs[id(dy.buf)] = dy.buf
            tape.grad_init[id(dy.buf)] = set()
            eng._conv_bwd(tape, op, x, dy, False, bias_grad_is_zero=True)

    def k_apply():
        L.call("mtb200_in_bwd_apply", g.data_ptr(), Cc, 0, y.ptr(), y.ldc, 0, g.data_ptr(), Cc, 0, d, B, y.nvox, Cc,
               L.ptr(y.xform), L.ptr(y.meanrstd), L.ptr(gamma), L.ptr(red), L.ptr(dg), L.ptr(db), L.stream_ptr())

    k_conv(); k_apply(); eng.begin_step()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        ts = []
        for _ in range(5):
            eng.begin_step()
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[2]

    def both(first_conv):
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        if first_conv:
            with torch.cuda.stream(s1):
                k_conv()
            with torch.cuda.stream(s2):
                k_apply()
        else:
            with torch.cuda.stream(s2):
                k_apply()
            with torch.cuda.stream(s1):
                k_conv()
        cur.wait_stream(s1); cur.wait_stream(s2)

    a = timed(k_conv)
    b = timed(k_apply)
    print("%s %s alone %.3f ms; in_bwd_apply alone %.3f ms; sum %.3f" % (what, dims, a, b, a + b))
    print("  both, conv launched first:  %.3f ms" % timed(lambda: both(True)))
    print("  both, apply launched first: %.3f ms" % timed(lambda: both(False)))


if __name__ == "__main__":
    main()
