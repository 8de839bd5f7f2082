#!/usr/bin/env python
"""Micro-benchmark of single convolution layers of the Generic_UNet stack (CUDA events, L2 flushed between reps).
usage: python tools/conv_bench.py [--impl 3 4] [--what fwd dgrad wgrad] [--layers all|full|l1|...]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch import nn  # noqa: E402
from multitalent_b200 import _lib as L  # noqa: E402
from multitalent_b200.engine import ConvOp, Engine, Feat, Tape  # noqa: E402

# (name, cin, cout, dims(B,D,H,W), split)
LAYERS = {
    "l0a": (1, 30, (4, 192, 160, 128), 0),
    "l0b": (30, 30, (4, 192, 160, 128), 0),
    "l0d": (60, 30, (4, 192, 160, 128), 30),
    "l1b": (60, 60, (4, 96, 80, 64), 0),
    "l1d": (120, 60, (4, 96, 80, 64), 60),
    "l2b": (120, 120, (4, 48, 40, 32), 0),
    "l2d": (240, 120, (4, 48, 40, 32), 120),
    "l3b": (240, 240, (4, 24, 20, 16), 0),
    "l3d": (480, 240, (4, 24, 20, 16), 240),
    "l4b": (320, 320, (4, 12, 10, 8), 0),
    "l4d": (640, 320, (4, 12, 10, 8), 320),
    "l5b": (320, 320, (4, 12, 5, 4), 0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", type=int, nargs="+", default=[3, 4, 5])
    ap.add_argument("--what", nargs="+", default=["fwd", "dgrad", "wgrad"])
    ap.add_argument("--layers", nargs="+", default=list(LAYERS))
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dt = torch.bfloat16
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    print("%-5s %-6s %4s %9s %9s   (Cin_p,Cout_p,dims)" % ("layer", "what", "impl", "ms", "TFLOP/s"))
    for name in a.layers:
        cin, cout, dims, split = LAYERS[name]
        B, D, H, W = dims
        conv = nn.Conv3d(cin, cout, 3, 1, 1, bias=True).cuda()
        op = ConvOp(conv.weight, conv.bias, (3, 3, 3), (1, 1, 1), split=split)
        flops = 2.0 * cin * cout * 27 * B * D * H * W
        for impl in a.impl:
            eng = Engine(dt, impl)
            if cin == 1 and eng.use_c1(op):  # compact single-channel input, as the network feeds the first layer
                x = Feat(torch.randn(B, D, H, W, 1, device="cuda").to(dt), 0, 1, 1)
            else:
                x = Feat(torch.randn(B, D, H, W, op.Cin_p, device="cuda").to(dt), 0, cin, op.Cin_p)
            dy = Feat(torch.randn(B, D, H, W, op.Cout_p, device="cuda").to(dt), 0, cout, op.Cout_p)
            for what in a.what:
                def run():
                    if what == "fwd":
                        eng.conv(op, x, out=dy, want_stats=True)
                    else:
                        tape = Tape()
                        tape.grad_bufs[id(dy.buf)] = dy.buf
                        tape.grad_init[id(dy.buf)] = set()
                        if what == "dgrad":
                            gx, have = tape.grad_feat(x)
                            grid = gx.dims[1:]
                            eng._conv_call(op.dgrad_taps, dy, op.packed(eng.wdtype, True), None, gx, grid, None, False,
                                           op.Cout_p, op.Cin_p)
                        else:
                            eng._conv_bwd(tape, op, x, dy, False, bias_grad_is_zero=True)
                try:
                    run()
                    torch.cuda.synchronize()
                except L.Mtb200Error as e:
                    print("%-5s %-6s %4d   unsupported: %s" % (name, what, impl, str(e)[:60]))
                    continue
                ts = []
                for _ in range(a.reps):
                    flush.zero_()
                    with L.KernelProfile() as kp:
                        run()
                    rows = kp.per_launch()
                    ts.append(sum(r[2] for r in rows if r[0] in ("mtb200_conv_taps", "conv_fwd", "conv_dgrad", "conv_wgrad")))
                ms = sorted(ts)[len(ts) // 2]
                print("%-5s %-6s %4d %9.3f %9.1f   (%d,%d,%s)" % (name, what, impl, ms, flops / ms / 1e9, op.Cin_p,
                                                                   op.Cout_p, dims))
            del x, dy
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
