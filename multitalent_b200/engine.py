"""Host-side executor for the native U-Net path: channels-last (NDHWC) feature buffers, tap tables for every
convolution variant, and a closure tape for the backward pass.  All arithmetic happens in libmtb200.so (ctypes);
PyTorch is used for device memory, streams and tiny bookkeeping tensors only.

Data layout (DESIGN.md section "HBM layout"): a feature map is a slice [coff, coff+Cp) of a buffer [B, D, H, W, ldc];
channel counts are padded (30->32, 60->64, 120->128, 240->256, 320->320, 47->48, 1->16) and padded channels are zero.
A conv writes its RAW output plus per-(b, c) sum / sum-of-squares; InstanceNorm + LeakyReLU are carried as a pending
per-(b, c) affine `xform` and applied by the CONSUMER while it loads the tile (norm-on-load), so the reference's
instnorm/lrelu read+write passes (generic_UNet.py:70) and torch.cat (generic_UNet.py:392) disappear.
"""
import contextlib
import ctypes as C
import os
import weakref
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

LRELU_SLOPE = 0.01
IN_EPS = 1e-5


_weights_epoch = 0


def bump_weights_epoch():
    """Call after parameters were modified behind torch's version counters (in-place arena update by a kernel)."""
    global _weights_epoch
    _weights_epoch += 1


def pad_channels(c: int) -> int:
    p = max(16, (c + 15) // 16 * 16)
    if p > 128 and p % 64:
        p = (p + 63) // 64 * 64
    return p


# --------------------------------------------------------------------------------------------------------------------
# tap tables
# --------------------------------------------------------------------------------------------------------------------
@dataclass
class TapTable:
    taps: List[Tuple[Tuple[int, int, int], int]]            # ((dz, dy, dx), weight slice)
    group_begin: List[int]
    group_ooff: List[Tuple[int, int, int]]
    in_stride: Tuple[int, int, int]
    out_stride: Tuple[int, int, int]

    def fill(self, p):
        assert len(self.taps) <= L.MAX_TAPS and len(self.group_ooff) <= L.MAX_GROUPS
        p.ntaps = len(self.taps)
        p.ngroups = len(self.group_ooff)
        for i, v in enumerate(self.group_begin):
            p.group_tap_begin[i] = v
        for g, o in enumerate(self.group_ooff):
            for k in range(3):
                p.group_ooff[g][k] = o[k]
        for t, (off, widx) in enumerate(self.taps):
            for k in range(3):
                p.tap_off[t][k] = off[k]
            p.tap_widx[t] = widx
        for k in range(3):
            p.is_[k] = self.in_stride[k]
            p.os_[k] = self.out_stride[k]


def _widx(k, kernel):
    return (k[0] * kernel[1] + k[1]) * kernel[2] + k[2]


def _kernel_positions(kernel):
    return [(a, b, c) for a in range(kernel[0]) for b in range(kernel[1]) for c in range(kernel[2])]


def taps_conv_fwd(kernel, stride) -> TapTable:
    """Conv3d(kernel, stride, padding=(k-1)//2): out[o] = sum_k W[k] in[o*s + k - p]."""
    pad = [(k - 1) // 2 for k in kernel]
    taps = [(tuple(k[i] - pad[i] for i in range(3)), _widx(k, kernel)) for k in _kernel_positions(kernel)]
    return TapTable(taps, [0, len(taps)], [(0, 0, 0)], tuple(stride), (1, 1, 1))


_WPAIR_TAPS = {}


def wpair_taps(stride) -> TapTable:
    """Conv3d(3x3x3, stride (sd, sh, 2)) on the [.., W/2, 2C] pair view of its input: output voxel q reads the pair q
    (taps kx = 1, 2) and the odd voxel of the pair q - 1 (kx = 0); the w stride of the view is 1."""
    key = tuple(stride)
    if key not in _WPAIR_TAPS:
        taps = []
        for kz in range(3):
            for ky in range(3):
                taps.append(((kz - 1, ky - 1, 0), (kz * 3 + ky) * 2))
                taps.append(((kz - 1, ky - 1, -1), (kz * 3 + ky) * 2 + 1))
        _WPAIR_TAPS[key] = TapTable(taps, [0, len(taps)], [(0, 0, 0)], (stride[0], stride[1], 1), (1, 1, 1))
    return _WPAIR_TAPS[key]


def wpair_weights(wp: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Weights of `wpair_taps`: [18][Cout][2C] from the packed forward layout [27][Cout][C] of a 3x3x3 convolution.
    Slice (kz, ky, 0) = [W(kx=1) | W(kx=2)] (the pair at the output's own position), slice (kz, ky, 1) = [0 | W(kx=0)]
    (the pair one to the left: only its odd voxel)."""
    Co, Ci = wp.shape[1], wp.shape[2]
    w4 = wp.view(3, 3, 3, Co, Ci)
    wv = out if out is not None else torch.zeros((3, 3, 2, Co, 2 * Ci), dtype=wp.dtype, device=wp.device)
    wv[:, :, 0, :, :Ci] = w4[:, :, 1]
    wv[:, :, 0, :, Ci:] = w4[:, :, 2]
    wv[:, :, 1, :, Ci:] = w4[:, :, 0]
    return wv


def wpair_weights_convT_dgrad(wd: torch.Tensor, kernel, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Weights of `wpair_taps_convT_dgrad`: [k0 k1][Cin][2 Cout] from the data-gradient layout [k0 k1 2][Cin][Cout] of a
    ConvTranspose3d(k == s): K = (kx, co)."""
    k0, k1 = kernel[0], kernel[1]
    Ci, Co = wd.shape[1], wd.shape[2]
    w4 = wd.view(k0, k1, 2, Ci, Co)
    wv = out if out is not None else torch.empty((k0, k1, Ci, 2 * Co), dtype=wd.dtype, device=wd.device)
    wv[..., :Co] = w4[:, :, 0]
    wv[..., Co:] = w4[:, :, 1]
    return wv


def wpair_taps_conv_dgrad(stride) -> TapTable:
    """Data gradient of Conv3d(3x3x3, stride (sd, sh, 2)) on the [.., W/2, 2C] pair view of d_in: the output pair
    (2q, 2q + 1) = the residues rw = 0, 1 of q, so only (rz, ry) remain as groups and the pair is 2C output channels
    (rw, ci).  W taps: dy[q] feeds rw = 0 through kx = 1 and rw = 1 through kx = 2; dy[q + 1] feeds rw = 1 through kx = 0."""
    key = ("D",) + tuple(stride)
    if key not in _WPAIR_TAPS:
        def dim_taps(r, s):
            return [(k, (r - k + 1) // s) for k in range(3) if (r - k + 1) % s == 0]
        taps, begin, ooff = [], [0], []
        for rz in range(stride[0]):
            for ry in range(stride[1]):
                for kz, oz in dim_taps(rz, stride[0]):
                    for ky, oy in dim_taps(ry, stride[1]):
                        for woff in (0, 1):
                            taps.append(((oz, oy, woff), (kz * 3 + ky) * 2 + woff))
                begin.append(len(taps))
                ooff.append((rz, ry, 0))
        _WPAIR_TAPS[key] = TapTable(taps, begin, ooff, (1, 1, 1), (stride[0], stride[1], 1))
    return _WPAIR_TAPS[key]


def wpair_weights_conv_dgrad(wd: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Weights of `wpair_taps_conv_dgrad`: [18][2 Cin][Cout] from the data-gradient layout [27][Cin][Cout]: slice
    (kz, ky, 0) = rows [W(kx=1); W(kx=2)], slice (kz, ky, 1) = rows [0; W(kx=0)]."""
    Ci, Co = wd.shape[1], wd.shape[2]
    w4 = wd.view(3, 3, 3, Ci, Co)
    wv = out if out is not None else torch.zeros((3, 3, 2, 2 * Ci, Co), dtype=wd.dtype, device=wd.device)
    wv[:, :, 0, :Ci] = w4[:, :, 1]
    wv[:, :, 0, Ci:] = w4[:, :, 2]
    wv[:, :, 1, Ci:] = w4[:, :, 0]
    return wv


def _pair_convT_taps(kernel) -> TapTable:
    """Forward tap table of the transposed convolution with kernel == stride == (k0, k1, 1) that a ConvTranspose3d with
    kernel (k0, k1, 2) is on the w-pair view of its output."""
    key = ("Tf",) + tuple(kernel)
    if key not in _WPAIR_TAPS:
        _WPAIR_TAPS[key] = taps_convT_fwd((kernel[0], kernel[1], 1))
    return _WPAIR_TAPS[key]


def wpair_taps_convT_dgrad(kernel) -> TapTable:
    """Data gradient of ConvTranspose3d(kernel == stride, kernel[2] == 2) on the [.., W/2, 2C] pair view of dy."""
    key = ("T",) + tuple(kernel)
    if key not in _WPAIR_TAPS:
        taps = [((kz, ky, 0), kz * kernel[1] + ky) for kz in range(kernel[0]) for ky in range(kernel[1])]
        _WPAIR_TAPS[key] = TapTable(taps, [0, len(taps)], [(0, 0, 0)], (kernel[0], kernel[1], 1), (1, 1, 1))
    return _WPAIR_TAPS[key]


def taps_conv_dgrad(kernel, stride) -> TapTable:
    """Data gradient of the above: d_in[s*q + r] = sum_{k: (r - k + p) % s == 0} W[k]^T dy[q + (r - k + p)/s].
    One group per residue class r (prod(stride) groups); every input voxel is written exactly once."""
    pad = [(k - 1) // 2 for k in kernel]
    taps, begin, ooff = [], [0], []
    for r in _kernel_positions(stride):
        for k in _kernel_positions(kernel):
            if all((r[i] - k[i] + pad[i]) % stride[i] == 0 for i in range(3)):
                off = tuple((r[i] - k[i] + pad[i]) // stride[i] for i in range(3))
                taps.append((off, _widx(k, kernel)))
        begin.append(len(taps))
        ooff.append(r)
    return TapTable(taps, begin, ooff, (1, 1, 1), tuple(stride))


def taps_convT_fwd(kernel) -> TapTable:
    """ConvTranspose3d(kernel == stride): out[q*s + k] = W[k] in[q]; one single-tap group per kernel position."""
    taps, begin, ooff = [], [0], []
    for k in _kernel_positions(kernel):
        taps.append(((0, 0, 0), _widx(k, kernel)))
        begin.append(len(taps))
        ooff.append(k)
    return TapTable(taps, begin, ooff, (1, 1, 1), tuple(kernel))


def taps_convT_dgrad(kernel) -> TapTable:
    """Data gradient of ConvTranspose3d(kernel == stride): d_in[q] = sum_k W[k]^T dy[q*s + k]."""
    taps = [(k, _widx(k, kernel)) for k in _kernel_positions(kernel)]
    return TapTable(taps, [0, len(taps)], [(0, 0, 0)], tuple(kernel), (1, 1, 1))


# --------------------------------------------------------------------------------------------------------------------
# feature maps
# --------------------------------------------------------------------------------------------------------------------
@dataclass
class Feat:
    buf: torch.Tensor                   # [B, D, H, W, ldc]
    coff: int
    C: int                              # logical channels
    Cp: int                             # padded channels (slice width)
    xform: Optional[torch.Tensor] = None     # [B, Cp, 4] pending {scale, shift, slope, 0}; None = already materialised
    meanrstd: Optional[torch.Tensor] = None  # [B, Cp, 2]
    act: Optional["Feat"] = None             # cached materialised activation f(raw) (tensor-core path)
    raw_of: Optional[Callable] = None        # on a materialised activation: weak reference to the raw conv output it was
    #                                          computed from (weak: raw.act -> activation must not close a cycle, or every
    #                                          step's buffers would wait for the cyclic garbage collector)
    single_consumer: bool = False            # exactly one layer consumes this output: its dgrad holds the FINAL d(act)
    red_fused: Optional[torch.Tensor] = None  # InstanceNorm-backward sums already accumulated by that dgrad's epilogue
    # PLANAR concatenation (full-resolution skip level of the tensor-core path): `planar` = the [2, B, D, H, W, Cp] tensor
    # whose halves are the up-sampled features (0) and the skip (1), each a compact NDHWC tensor.  `half` = 0 / 1: this
    # Feat is one half (buf = planar[half], an ordinary compact feature); None: the concatenation of both (buf = the
    # [2B, D, H, W, Cp] view; only the line-streaming conv kernels read / write it, mtb200_conv_params::in_split).
    planar: Optional[torch.Tensor] = None
    half: Optional[int] = None

    @property
    def dims(self):
        if self.planar is not None and self.half is None:
            return (self.buf.shape[0] // 2,) + tuple(self.buf.shape[1:4])
        return tuple(self.buf.shape[:4])

    @property
    def split(self):
        """Channels per planar half when this Feat is a planar concatenation, else 0."""
        return self.buf.shape[4] if (self.planar is not None and self.half is None) else 0

    @property
    def ldc(self):
        return self.buf.shape[4]

    @property
    def nvox(self):
        return self.buf.shape[1] * self.buf.shape[2] * self.buf.shape[3]

    def ptr(self):
        return self.buf.data_ptr()

    def as_ncdhw(self):
        """NCDHW-shaped strided view of the logical channels (what the reference's callers index)."""
        return self.buf[..., self.coff:self.coff + self.C].permute(0, 4, 1, 2, 3)


def ndhwc_view_info(t: torch.Tensor):
    """If `t` (NCDHW-shaped) is a channels-last view of an NDHWC buffer, return (ldc); else None."""
    if t.dim() != 5:
        return None
    B, Cc, D, H, W = t.shape
    s = t.stride()
    if s[1] != 1:
        return None
    ldc = s[4]
    if ldc < Cc or ldc % 8 or s[3] != W * ldc or s[2] != H * W * ldc or s[0] != D * H * W * ldc:
        return None
    if t.data_ptr() % 16:
        return None
    return ldc


class ConvOp:
    """One convolution's parameters + packed copies.  `weight` is the reference-layout nn.Parameter
    ([Cout, Cin, kd, kh, kw], or [Cin, Cout, kd, kh, kw] for ConvTranspose3d)."""

    def __init__(self, weight, bias, kernel, stride, transposed=False, split=0):
        self.weight, self.bias = weight, bias
        self.kernel, self.stride = tuple(int(k) for k in kernel), tuple(int(s) for s in stride)
        self.transposed = transposed
        if transposed:
            self.Cin, self.Cout = weight.shape[0], weight.shape[1]
            assert self.kernel == self.stride, "only ConvTranspose3d with kernel == stride is on the native path"
        else:
            self.Cout, self.Cin = weight.shape[0], weight.shape[1]
        self.split = split  # logical channels of the first half of a concatenated input (0 = plain input)
        if split:
            self.split_p = pad_channels(split)
            self.Cin_p = self.split_p + pad_channels(self.Cin - split)
        else:
            self.split_p = 0
            self.Cin_p = pad_channels(self.Cin)
        self.Cout_p = pad_channels(self.Cout)
        self.ntap = self.kernel[0] * self.kernel[1] * self.kernel[2]
        if transposed:
            self.fwd_taps = taps_convT_fwd(self.kernel)
            self.dgrad_taps = taps_convT_dgrad(self.kernel)
        else:
            self.fwd_taps = taps_conv_fwd(self.kernel, self.stride)
            self.dgrad_taps = taps_conv_dgrad(self.kernel, self.stride)
        self._packed = {}
        self.group = None  # PackGroup: every op of the network refreshed by one launch
        # first layer on a single-channel patch: dedicated kernels with K = taps (csrc/conv_c1.cu)
        self.c1 = (not transposed and self.Cin == 1 and self.kernel == (3, 3, 3) and self.stride == (1, 1, 1)
                   and split == 0 and self.Cout_p in (32, 64))

    def out_dims(self, dims):
        B, D, H, W = dims
        if self.transposed:
            return (B, D * self.stride[0], H * self.stride[1], W * self.stride[2])
        for n, s in zip((D, H, W), self.stride):
            assert n % s == 0, "spatial size %s not divisible by stride %s" % ((D, H, W), self.stride)
        return (B, D // self.stride[0], H // self.stride[1], W // self.stride[2])

    def packed(self, wdtype, swap_io):
        """[ntap][Cout_p][Cin_p] (forward / wgrad layout) or, swap_io, [ntap][Cin_p][Cout_p] (dgrad operand)."""
        key = (wdtype, swap_io)
        ver = (self.weight._version, _weights_epoch, self.weight.data_ptr())
        hit = self._packed.get(key)
        if hit is not None and hit[0] == ver and hit[1].device == self.weight.device:
            return hit[1]
        if self.group is not None and self.group.refresh(wdtype):
            return self._packed[key][1]
        w = self.weight.detach()
        assert w.dtype == torch.float32 and w.is_contiguous()
        out = torch.empty((self.ntap, self.Cin_p, self.Cout_p) if swap_io else (self.ntap, self.Cout_p, self.Cin_p),
                          dtype=wdtype, device=w.device)
        L.call("mtb200_pack_weights", L.ptr(w), self.Cout, self.Cin, self.ntap, int(self.transposed), int(swap_io),
               L.ptr(out), L.dtype_enum(wdtype), self.Cout_p, self.Cin_p, self.split, self.split_p, L.stream_ptr())
        self._packed[key] = (ver, out)
        return out


class PackGroup:
    """The convolutions of one network.  The optimizer step invalidates every packed weight copy at once, so the copies
    of ALL ops (forward layout and the data-gradient layout) are rebuilt by ONE launch (`mtb200_pack_weights_batched`)
    the first time any of them is asked for after a step, instead of ~60 per-layer launches."""

    def __init__(self, ops: Sequence["ConvOp"]):
        self.ops = list(ops)
        for op in self.ops:
            op.group = self
        self._state = {}  # wdtype -> (descriptor table on the device, total blocks, [(packed, packed_swap)], pointers)

    def _build(self, wdtype):
        dev = self.ops[0].weight.device
        descs = (L.PackDesc * len(self.ops))()
        bufs, blk = [], 0
        for d, op in zip(descs, self.ops):
            w = op.weight.detach()
            a = torch.empty((op.ntap, op.Cout_p, op.Cin_p), dtype=wdtype, device=dev)
            b = torch.empty((op.ntap, op.Cin_p, op.Cout_p), dtype=wdtype, device=dev)
            d.w, d.packed, d.packed_swap = w.data_ptr(), a.data_ptr(), b.data_ptr()
            d.Cout, d.Cin, d.ntap, d.transposed = op.Cout, op.Cin, op.ntap, int(op.transposed)
            d.Cout_p, d.Cin_p, d.split, d.split_p = op.Cout_p, op.Cin_p, op.split, op.split_p
            d.blk_begin = blk
            blk += (op.Cout_p // 16) * (op.Cin_p // 16)
            bufs.append((a, b))
        table = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
        ptrs = tuple(op.weight.data_ptr() for op in self.ops)
        self._state[wdtype] = (table, blk, bufs, ptrs)

    def refresh(self, wdtype) -> bool:
        """Repack every op for `wdtype`; False if the group cannot serve the request (caller packs the single op)."""
        ops = self.ops
        w0 = ops[0].weight
        if not w0.is_cuda or any(op.ntap > 32 or op.weight.dtype != torch.float32 or not op.weight.is_contiguous()
                                 or op.weight.device != w0.device for op in ops):
            return False
        st = self._state.get(wdtype)
        if st is None or st[3] != tuple(op.weight.data_ptr() for op in ops) or st[0].device != w0.device:
            self._build(wdtype)
            st = self._state[wdtype]
        table, blocks, bufs, _ = st
        L.call("mtb200_pack_weights_batched", L.ptr(table), len(ops), blocks, L.dtype_enum(wdtype), L.stream_ptr(),
               tag="mtb200_pack_weights")
        for op, (a, b) in zip(ops, bufs):
            ver = (op.weight._version, _weights_epoch, op.weight.data_ptr())
            op._packed[(wdtype, False)] = (ver, a)
            op._packed[(wdtype, True)] = (ver, b)
        return True


def collect_ops(tree) -> List["ConvOp"]:
    """Every ConvOp inside nested tuples / lists / dicts, in traversal order."""
    if isinstance(tree, ConvOp):
        return [tree]
    if isinstance(tree, dict):
        tree = list(tree.values())
    if isinstance(tree, (list, tuple)):
        return [op for t in tree for op in collect_ops(t)]
    return []


def _padded(v: Optional[torch.Tensor], n: int, fill=0.0):
    """1-D parameter zero-padded to `n` entries.  Parameters that live in a FlatArena own a padded slot (the tail stays
    zero under SGD): they are viewed in place, no copy."""
    if v is None:
        return None
    if getattr(v, "_mtb_padded_len", 0) >= n and v.dtype == torch.float32:
        return v.detach().as_strided((n,), (1,))
    v = v.detach().float()
    if v.numel() == n:
        return v.contiguous()
    out = torch.full((n,), fill, dtype=torch.float32, device=v.device)
    out[:v.numel()] = v
    return out


def direct_grad(p, n: Optional[int] = None):
    """The gradient slot of an arena parameter (FlatArena sets `_mtb_direct_grad`): kernels accumulate into it in place
    and autograd is handed no tensor for it.  `n`: padded length for 1-D parameters.  None = go through autograd."""
    if p is None or not getattr(p, "_mtb_direct_grad", False) or p.grad is None:
        return None
    if n is None:
        return p.grad
    if getattr(p, "_mtb_padded_len", 0) >= n:
        return p.grad.as_strided((n,), (1,))
    return None


class ZeroPool:
    """Bump allocator over zero-filled device memory for the per-step accumulation buffers (statistics, reductions,
    packed weight gradients): one memset per step instead of ~80 torch.zeros launches."""

    def __init__(self, dtype, chunk_elems):
        self.dtype, self.chunk_elems = dtype, chunk_elems
        self.chunks = []  # [tensor, used]

    def reset(self):
        for c in self.chunks:
            if c[1]:
                c[0][:c[1]].zero_()
                c[1] = 0

    def take(self, shape, device):
        n = 1
        for d in shape:
            n *= int(d)
        n16 = (n + 15) // 16 * 16
        for c in self.chunks:
            if c[0].device == device and c[0].numel() - c[1] >= n16:
                t = c[0][c[1]:c[1] + n].view(shape)
                c[1] += n16
                return t
        c = [torch.zeros(max(n16, self.chunk_elems), dtype=self.dtype, device=device), n16]
        self.chunks.append(c)
        return c[0][:n].view(shape)


class Tape:
    """Records what backward needs.  Gradients w.r.t. feature buffers are kept per underlying buffer so that the two
    halves of a concatenated skip buffer share one gradient buffer."""

    def __init__(self):
        self.closures: List[Callable] = []
        self.grad_bufs: Dict[int, torch.Tensor] = {}
        self.grad_init: Dict[int, set] = {}
        self.param_grads: Dict[int, torch.Tensor] = {}
        self.direct_done = set()  # ids of parameters whose gradient went straight into their arena slot
        self.keep = []  # keep python references to buffers alive

    @staticmethod
    def _key_blocks(f: Feat):
        """(gradient-buffer key, 16-channel blocks of it that `f` covers).  The halves and the concatenation of a planar
        buffer share ONE gradient buffer of the same planar layout (keyed by the planar tensor)."""
        if f.planar is not None:
            n16 = f.planar.shape[5] // 16
            h0, h1 = (0, 2) if f.half is None else (f.half, f.half + 1)
            return id(f.planar), set(range(h0 * n16, h1 * n16))
        return id(f.buf), set(range(f.coff // 16, (f.coff + f.Cp) // 16))

    def grad_feat(self, f: Feat) -> Tuple[Feat, bool]:
        """Gradient slice matching `f` and whether it already holds a value (=> producers must accumulate)."""
        k, blocks = self._key_blocks(f)
        if k not in self.grad_bufs:
            self.grad_bufs[k] = torch.empty_like(f.planar if f.planar is not None else f.buf)
            self.grad_init[k] = set()
        have = blocks & self.grad_init[k]
        assert not have or have == blocks, "partially initialised gradient slice"
        gb = self.grad_bufs[k]
        if f.planar is not None:
            buf = gb.view((-1,) + tuple(gb.shape[2:])) if f.half is None else gb[f.half]
            return Feat(buf, 0, f.C, f.Cp, planar=gb, half=f.half), bool(have)
        return Feat(gb, f.coff, f.C, f.Cp), bool(have)

    def mark(self, f: Feat):
        k, blocks = self._key_blocks(f)
        self.grad_init[k] |= blocks

    def has_grad(self, f: Feat) -> bool:
        k, blocks = self._key_blocks(f)
        return k in self.grad_init and blocks <= self.grad_init[k]

    def add_param_grad(self, p, g):
        if p is None:
            return
        k = id(p)
        if k in self.param_grads:
            self.param_grads[k] = self.param_grads[k] + g
        else:
            self.param_grads[k] = g


class Engine:
    """Kernel-call layer.  `dtype` is the activation storage type: torch.float32 (T0 parity mode, CUDA-core kernels),
    torch.bfloat16 / torch.float16 (T1, tensor-core kernels where available)."""

    def __init__(self, dtype=torch.float32, impl=0):
        self.dtype = dtype
        self.impl = impl  # 0 auto, 1 force FFMA, 2 force tcgen05 (3..7: one tcgen05 kernel family only, see mtb200.h; 8: as 2)
        self.wdtype = torch.float32 if dtype == torch.float32 else dtype
        self._materialize = None
        self._z64 = ZeroPool(torch.float64, 1 << 16)
        self._z32 = ZeroPool(torch.float32, 1 << 22)
        # weight gradients are only consumed by the optimizer: with arena parameters (gradients accumulated in place, no
        # autograd tensors) they run on a second stream next to the dgrad / InstanceNorm-backward chain, which fills the
        # SMs the small deep-level launches and every kernel's tail leave idle.  MTB200_WGRAD_STREAM=0 disables it.
        self.overlap_wgrad = os.environ.get("MTB200_WGRAD_STREAM", "1") != "0"
        self._side = {}
        self._side_used = None
        # arena mode: the packed fp32 weight gradients of a step are folded into the parameters' gradient slots by ONE
        # launch at the end of the backward pass (MTB200_UNPACK_BATCHED=0: one launch per layer, as before)
        self.batch_unpack = os.environ.get("MTB200_UNPACK_BATCHED", "1") != "0"
        # InstanceNorm-backward reduction fused into the consumer's data-gradient epilogue where the kernel supports it
        self.fuse_red = os.environ.get("MTB200_FUSE_RED", "1") != "0"
        # Heads under the MultiTalent loss: loss pass 2 + head data gradient + head weight gradient as ONE kernel
        # (mtb200_head_bwd_fused).  The loss's backward leaves a descriptor in `lazy_heads` instead of writing d(logits).
        self.fuse_head = os.environ.get("MTB200_FUSE_HEAD", "1") != "0"
        self.fusable_heads = {}  # logits buffer pointer -> True, for the heads of the current step the kernel covers
        self.lazy_heads = {}     # logits buffer pointer -> what the fused kernel needs from the loss
        # Deferred heads (set by the trainer around the network call of a training step, never by default): a fusable head
        # is not computed at all in the forward pass -- the loss's first pass evaluates the logits window on the tensor core
        # from the head's input (mtb200_head_fwd_stats) and the fused backward recomputes it.  The network then returns a
        # zero-stride placeholder for that output; only multitalent_loss(engine=...) may consume it.
        self.planar_concat = os.environ.get("MTB200_PLANAR", "1") != "0"
        # stride-2 forward on the w-pair view of a dense 32-channel input (Engine.conv / wpair_taps)
        self.wpairs = os.environ.get("MTB200_WPAIRS", "1") != "0"
        self.wpairs_dgrad = os.environ.get("MTB200_WPAIRS_DGRAD", "1") != "0"
        self.defer_head_fwd = os.environ.get("MTB200_DEFER_HEAD", "1") != "0"
        self.defer_heads = False
        self.deferred = {}       # placeholder pointer -> {"x": head input, "op": head}
        # Inference: the sliding-window predictor asks for the head's INPUT instead of the logits (`capture_head`) and runs
        # head -> sigmoid x Gaussian -> scatter-add as one kernel per tile (mtb200_head_aggregate)
        self.fuse_head_aggregate = os.environ.get("MTB200_FUSE_HEAD_AGG", "1") != "0"
        self.capture_head = False
        self.captured_head = None
        self.wgrad_order = int(os.environ.get("MTB200_WGRAD_ORDER", "0"))
        self.bwd_priority = os.environ.get("MTB200_BWD_PRIO", "0") != "0"
        self._hp = {}
        self._pending_unpack = []
        self._unpack_tables = {}
        # data parallel: called once in the backward pass, when every gradient except those of the first encoder stages is
        # final (set by the trainer; the network places the call, see Generic_UNet._native_forward)
        self.backward_mark: Optional[Callable] = None
        self.apply_inplace = os.environ.get("MTB200_APPLY_INPLACE", "1") != "0"
        self._dy_bufs = []   # d(raw output) buffers when not in place: reused from step to step in call order
        self._dy_next = 0

    def begin_step(self):
        """Called at the start of every network forward: re-zero what the previous step took from the pools."""
        self._z64.reset()
        self._z32.reset()
        self._dy_next = 0
        self.fusable_heads.clear()
        self.lazy_heads.clear()
        self.deferred.clear()

    @property
    def materialize_inputs(self) -> bool:
        """The tcgen05 conv kernel feeds TMA tiles straight to the tensor core, so its inputs must already be
        normalised + activated; the CUDA-core kernels apply the pending transform while loading instead."""
        if self._materialize is None:
            self._materialize = bool(self.dtype != torch.float32 and self.impl != 1
                                     and L.lib().mtb200_has_tcgen05() == 1)
        return self._materialize

    def operand(self, x: Feat) -> Feat:
        """The tensor a conv kernel reads for `x`: `x` itself (norm-on-load) or its cached materialised activation."""
        if x.xform is None or not self.materialize_inputs:
            return x
        if x.act is None:
            x.act = self.materialize(x)
        return x.act

    # ---- helpers ------------------------------------------------------------------------------------------------
    def new_buf(self, dims, ldc, device, zero=False):
        f = torch.zeros if zero else torch.empty
        return f(tuple(dims) + (ldc,), dtype=self.dtype, device=device)

    def planar_concat_ok(self, half_Cp: int, dims) -> bool:
        """Planar skip / up-sampling halves for a decoder level (instead of interleaved channels in one buffer)?  Where a
        half is 32 channels = 64 bytes per voxel, every kernel that touches ONE half of the interleaved buffer (strided
        conv of the next encoder stage, InstanceNorm passes of the skip, transposed conv and its gradients) moved whole
        128-byte lines for 64 useful bytes.  Needs the line-streaming kernels for the level's first decoder conv."""
        _, D, H, W = dims
        return bool(self.planar_concat and self.materialize_inputs and self.impl in (0, 2, 8) and half_Cp == 32
                    and W >= 72 and H >= 4)

    def use_c1(self, op: "ConvOp") -> bool:
        """First-layer kernels (K = taps) apply: single input channel, 16-bit tensor-core path."""
        return bool(op.c1 and self.materialize_inputs and self.impl in (0, 2, 8))

    def input_feat(self, x: torch.Tensor, compact=False) -> Feat:
        """NCDHW fp32 input (to_torch.py:18-31 contract) -> NDHWC materialised feature (zero-padded channels).
        `compact` (single-channel input feeding the first-layer kernels): no channel padding, [B, D, H, W, 1]."""
        assert x.dim() == 5 and x.is_cuda, "input must be a [B, C, D, H, W] CUDA tensor"
        if compact and x.shape[1] == 1:
            B, _, D, H, W = x.shape
            return Feat(x.detach().to(self.dtype).contiguous().view(B, D, H, W, 1), 0, 1, 1)
        x = x.detach().float().contiguous()
        B, Cc, D, H, W = x.shape
        Cp = pad_channels(Cc)
        buf = self.new_buf((B, D, H, W), Cp, x.device)
        L.call("mtb200_ncdhw_to_ndhwc", L.ptr(x), B, Cc, D * H * W, L.ptr(buf), L.dtype_enum(self.dtype), Cp, 0, Cp,
               L.stream_ptr())
        return Feat(buf, 0, Cc, Cp)

    @staticmethod
    def conv_flops(op: "ConvOp", out_dims):
        """Algorithmic FLOPs (2 x MAC, true channel counts) of one convolution; dgrad and wgrad cost the same."""
        n = out_dims[0] * out_dims[1] * out_dims[2] * out_dims[3]
        return 2.0 * op.Cin * op.Cout * n * (1 if op.transposed else op.ntap)

    def _conv_call(self, table: TapTable, x: Feat, w_packed, bias_p, out: Feat, grid_dims, stats, accumulate, Cin_p,
                   Cout_p, flops=0.0, tag="conv_fwd", red=None):
        """`red` = (raw Feat of the layer whose d(activation) this launch writes, fp64 [B, Cp, 2] accumulator): offer the
        fused InstanceNorm-backward reduction to the kernel.  Returns True if the dispatched kernel took it."""
        p = L.ConvParams()
        p.inp, p.out, p.w = x.ptr(), out.ptr(), w_packed.data_ptr()
        p.bias = bias_p.data_ptr() if bias_p is not None else None
        p.xform = x.xform.data_ptr() if x.xform is not None else None
        p.stats = stats.data_ptr() if stats is not None else None
        p.dtype, p.wdtype = L.dtype_enum(self.dtype), L.dtype_enum(w_packed.dtype)
        p.B, p.Di, p.Hi, p.Wi = x.dims
        p.in_ldc, p.in_coff, p.Cin = x.ldc, x.coff, Cin_p
        _, p.Dof, p.Hof, p.Wof = out.dims
        p.out_ldc, p.out_coff, p.Cout = out.ldc, out.coff, Cout_p
        p.in_split, p.out_split = x.split, out.split  # planar concatenations (line-streaming kernels only)
        p.Do, p.Ho, p.Wo = grid_dims
        table.fill(p)
        p.accumulate = int(accumulate)
        p.impl = self.impl
        if red is not None:
            raw, acc = red
            p.red_y, p.red_ldc, p.red_coff = raw.ptr(), raw.ldc, raw.coff
            p.red_xform, p.red_meanrstd, p.red = raw.xform.data_ptr(), raw.meanrstd.data_ptr(), acc.data_ptr()
        L.call("mtb200_conv_taps", C.byref(p), L.stream_ptr(), flops=flops, tag=tag,
               info=(Cin_p, Cout_p, tuple(grid_dims), len(table.taps), table.in_stride, table.out_stride))
        return red is not None and (L.lib().mtb200_last_kernel() or b"").endswith(b"+red")

    def _dy_buffer(self, dims, Cp, dev):
        """The i-th d(raw output) buffer of this step (MTB200_APPLY_INPLACE=0): persistent across steps -- the weight-gradient
        side streams that read it are joined before the next step starts -- and re-created when the shape changes."""
        i = self._dy_next
        self._dy_next += 1
        shape = tuple(dims) + (Cp,)
        if i < len(self._dy_bufs):
            t = self._dy_bufs[i]
            if tuple(t.shape) == shape and t.dtype == self.dtype and t.device == dev:
                return t
            self._dy_bufs[i] = t = torch.empty(shape, dtype=self.dtype, device=dev)
            return t
        t = torch.empty(shape, dtype=self.dtype, device=dev)
        self._dy_bufs.append(t)
        return t

    # ---- forward primitives -------------------------------------------------------------------------------------
    def conv(self, op: ConvOp, x: Feat, out: Optional[Feat] = None, want_stats=False):
        """Raw convolution (or transposed convolution) of `x` into `out` (allocated if None).  Returns (out, stats)."""
        c1 = self.use_c1(op) and x.xform is None
        assert x.Cp == op.Cin_p or (c1 and x.Cp == 1), "input slice width %d != conv's padded Cin %d" % (x.Cp, op.Cin_p)
        x = self.operand(x)
        odims = op.out_dims(x.dims)
        dev = x.buf.device
        if out is None:
            out = Feat(self.new_buf(odims, op.Cout_p, dev), 0, op.Cout, op.Cout_p)
        assert out.dims == odims and out.Cp == op.Cout_p
        stats = self._z64.take((odims[0], op.Cout_p, 2), dev) if want_stats else None
        if c1:
            B, D, H, W = x.dims
            bias = _padded(op.bias, op.Cout_p)
            L.call("mtb200_conv_c1_fwd", C.c_void_p(x.ptr() + x.coff * x.buf.element_size()), x.ldc,
                   L.ptr(op.packed(self.wdtype, False)), op.Cin_p, L.ptr(bias), out.ptr(), out.ldc, out.coff, op.Cout_p,
                   L.ptr(stats), L.dtype_enum(self.dtype), B, D, H, W, L.stream_ptr(),
                   flops=self.conv_flops(op, odims), tag="conv_fwd",
                   info=(1, op.Cout_p, tuple(odims[1:]), op.ntap, (1, 1, 1), (1, 1, 1)))
            return out, stats
        grid = x.dims[1:] if op.transposed else odims[1:]
        if self._use_wpairs(op, x):
            # stride 2 along w on a dense 32-channel tensor: voxels (2q, 2q + 1) are ONE 128-byte row of the
            # [B, D, H, W/2, 64] view of the same buffer -- 18 taps of 64 channels on dense rows instead of 27 taps of 32
            # channels on every other 64-byte row (see wpair_taps)
            B, D, H, W = x.dims
            xv = Feat(x.buf.view(B, D, H, W // 2, 2 * x.ldc), 0, 2 * x.ldc, 2 * x.ldc)
            self._conv_call(wpair_taps(op.stride), xv, self._wpair_weights(op), _padded(op.bias, op.Cout_p), out, grid,
                            stats, False, 2 * op.Cin_p, op.Cout_p, flops=self.conv_flops(op, odims), tag="conv_fwd")
            return out, stats
        if (op.bias is None and stats is None and self._wpairs_convT(op, out) and out.xform is None):
            # ConvTranspose3d(k == s, kx = 2) with 32 output channels on the w-pair view of its OUTPUT: kernel (k0, k1, 1),
            # 64 output channels (kx, co) -- the packed weights [(kz, ky, kx)][32][Cin] are [(kz, ky)][64][Cin] as they lie
            B, D, H, W = out.dims
            ov = Feat(out.buf.view(B, D, H, W // 2, 64), 0, 64, 64)
            wv = op.packed(self.wdtype, False).view(op.ntap // 2, 64, op.Cin_p)
            self._conv_call(_pair_convT_taps(op.kernel), x, wv, None, ov, grid, None, False, op.Cin_p, 64,
                            flops=self.conv_flops(op, odims), tag="conv_fwd")
            return out, stats
        self._conv_call(op.fwd_taps, x, op.packed(self.wdtype, False), _padded(op.bias, op.Cout_p), out, grid, stats,
                        False, op.Cin_p, op.Cout_p, flops=self.conv_flops(op, odims), tag="conv_fwd")
        return out, stats

    def _use_wpairs(self, op: ConvOp, x: Feat) -> bool:
        return (self.wpairs and not op.transposed and op.kernel == (3, 3, 3) and op.stride[2] == 2 and op.Cin_p == 32
                and x.ldc == 32 and x.coff == 0 and (x.planar is None or x.half is not None) and x.xform is None and x.dims[3] % 2 == 0
                and x.buf.is_contiguous() and self.dtype in (torch.bfloat16, torch.float16) and self.impl in (0, 2, 3))

    def _wpairs_convT(self, op: ConvOp, dy: Feat) -> bool:
        return (self.wpairs and op.transposed and op.stride[2] == 2 and op.Cout_p == 32 and dy.ldc == 32 and dy.coff == 0
                and dy.dims[3] % 2 == 0 and dy.buf.is_contiguous() and (dy.planar is None or dy.half is not None)
                and self.dtype in (torch.bfloat16, torch.float16) and self.impl in (0, 2, 3))

    def _wpair_weights_convT(self, op: ConvOp):
        """[k0 k1][Cin_p][64] from the data-gradient layout [k0 k1 2][Cin_p][32]: K = (kx, co)."""
        wd = op.packed(self.wdtype, True)
        ver = (op.weight._version, _weights_epoch, op.weight.data_ptr(), wd.data_ptr())
        hit = op._packed.get("wpairT")
        if hit is not None and hit[0] == ver:
            return hit[1]
        wv = wpair_weights_convT_dgrad(wd, op.kernel, hit[1] if hit is not None else None)
        op._packed["wpairT"] = (ver, wv)
        return wv

    def _wpair_weights(self, op: ConvOp):
        """[18][Cout_p][64] from the packed [27][Cout_p][32]: slice (kz, ky, 0) = [W(kx=1) | W(kx=2)] (the pair at the
        output's own position), slice (kz, ky, 1) = [0 | W(kx=0)] (the pair one to the left: only its odd voxel)."""
        wp = op.packed(self.wdtype, False)
        ver = (op.weight._version, _weights_epoch, op.weight.data_ptr(), wp.data_ptr())
        hit = op._packed.get("wpair")
        if hit is not None and hit[0] == ver:
            return hit[1]
        wv = wpair_weights(wp, hit[1] if hit is not None else None)
        op._packed["wpair"] = (ver, wv)
        return wv

    def finalize_norm(self, y: Feat, stats, gamma, beta, slope=LRELU_SLOPE):
        B = y.dims[0]
        dev = y.buf.device
        y.xform = torch.empty((B, y.Cp, 4), dtype=torch.float32, device=dev)
        y.meanrstd = torch.empty((B, y.Cp, 2), dtype=torch.float32, device=dev)
        L.call("mtb200_in_finalize", L.ptr(stats), L.ptr(gamma), L.ptr(beta), B, y.Cp, y.nvox, IN_EPS, slope,
               L.ptr(y.xform), L.ptr(y.meanrstd), L.stream_ptr())

    def materialize(self, x: Feat, res: Optional[Feat] = None, slope2=LRELU_SLOPE, out: Optional[Feat] = None) -> Feat:
        """act = f(x) [+ g(res), then LeakyReLU(slope2)] written to a fresh buffer (or `out`)."""
        if out is None:
            out = Feat(self.new_buf(x.dims, x.Cp, x.buf.device), 0, x.C, x.Cp)
        B = x.dims[0]
        if res is None:
            out.raw_of = weakref.ref(x)
        L.call("mtb200_norm_act", x.ptr(), x.ldc, x.coff, out.ptr(), out.ldc, out.coff, L.dtype_enum(self.dtype), B,
               x.nvox, x.Cp, L.ptr(x.xform), res.ptr() if res is not None else None,
               res.ldc if res is not None else 0, res.coff if res is not None else 0,
               L.ptr(res.xform) if res is not None else None, float(slope2), L.stream_ptr(),
               nbytes=(2 if res is None else 3) * B * x.nvox * x.Cp * x.buf.element_size(),
               info=("norm_act", x.Cp, x.ldc, out.ldc, tuple(x.dims[1:])))
        return out

    # ---- composite layers with tape ------------------------------------------------------------------------------
    def conv_norm(self, tape: Optional[Tape], op: ConvOp, gamma_param, beta_param, x: Feat, out: Optional[Feat] = None,
                  slope=LRELU_SLOPE, need_input_grad=True) -> Feat:
        """conv -> InstanceNorm (-> LeakyReLU(slope)); the norm/activation stay pending on the returned Feat."""
        y, stats = self.conv(op, x, out, want_stats=True)
        gamma = _padded(gamma_param, op.Cout_p)
        beta = _padded(beta_param, op.Cout_p)
        self.finalize_norm(y, stats, gamma, beta, slope)
        if tape is not None:
            tape.closures.append(lambda: self._conv_norm_bwd(tape, op, gamma_param, beta_param, gamma, x, y,
                                                             need_input_grad))
        return y

    def _conv_norm_bwd(self, tape, op, gamma_param, beta_param, gamma, x: Feat, y: Feat, need_input_grad):
        dev = y.buf.device
        B = y.dims[0]
        src = y.act if y.act is not None else y  # consumers of a materialised activation left their gradient there
        if not tape.has_grad(src):  # output never used downstream: all gradients are exactly zero
            self._zero_param_grads(tape, op, gamma_param, beta_param)
            return
        g, _ = tape.grad_feat(src)
        dt = L.dtype_enum(self.dtype)
        red, y.red_fused = y.red_fused, None
        if red is None:  # the consumer's data-gradient kernel did not accumulate the sums: separate pass over g and y
            red = self._z64.take((B, y.Cp, 2), dev)
            L.call("mtb200_in_bwd_reduce", g.ptr(), g.ldc, g.coff, y.ptr(), y.ldc, y.coff, dt, B, y.nvox, y.Cp,
                   L.ptr(y.xform), L.ptr(y.meanrstd), L.ptr(red), L.stream_ptr(),
                   nbytes=2 * B * y.nvox * y.Cp * y.buf.element_size(),
                   info=("in_bwd_reduce", y.Cp, g.ldc, y.ldc, tuple(y.dims[1:])))
        dgamma, dbeta = direct_grad(gamma_param, y.Cp), direct_grad(beta_param, y.Cp)
        direct = dgamma is not None and dbeta is not None
        if not direct:
            dgamma = torch.zeros(y.Cp, dtype=torch.float32, device=dev)
            dbeta = torch.zeros(y.Cp, dtype=torch.float32, device=dev)
        # In place by default (dy overwrites d(act); the kernel reads an aliased operand with coherent loads).
        # MTB200_APPLY_INPLACE=0: a persistent compact buffer per layer instead, both inputs on the read-only load path.
        if self.apply_inplace:
            dy = g
        else:
            dy = Feat(self._dy_buffer(y.dims, y.Cp, dev), 0, y.C, y.Cp)
        L.call("mtb200_in_bwd_apply", g.ptr(), g.ldc, g.coff, y.ptr(), y.ldc, y.coff, dy.ptr(), dy.ldc, dy.coff, dt, B,
               y.nvox, y.Cp, L.ptr(y.xform), L.ptr(y.meanrstd), L.ptr(gamma), L.ptr(red), L.ptr(dgamma), L.ptr(dbeta),
               L.stream_ptr(), nbytes=3 * B * y.nvox * y.Cp * y.buf.element_size(),
               info=("in_bwd_apply", y.Cp, g.ldc, y.ldc, tuple(y.dims[1:])))
        g = dy
        if direct:
            tape.direct_done.update((id(gamma_param), id(beta_param)))
        else:
            tape.add_param_grad(gamma_param, dgamma[:op.Cout])
            tape.add_param_grad(beta_param, dbeta[:op.Cout])
        # A bias in front of an InstanceNorm has an identically-zero gradient: sum_v dy = k1 (S1 - N m1 - m2 sum xhat)
        # = 0.  (The reference's autograd produces rounding noise ~1e-9 there.)  Emit exact zeros, skip the reduction.
        self._conv_bwd(tape, op, x, g, need_input_grad, bias_grad_is_zero=True)

    def _zero_param_grads(self, tape, op, *others):
        for p in (op.weight, op.bias) + others:
            if p is None:
                continue
            if direct_grad(p) is not None:
                tape.direct_done.add(id(p))  # the arena slot already holds this step's (zero) contribution
            else:
                tape.add_param_grad(p, torch.zeros_like(p))

    @contextlib.contextmanager
    def _wgrad_stream(self, op: ConvOp, dev, after=None):
        """Side stream for one layer's weight-gradient launches (ordered after everything already queued on the current
        stream, i.e. after d(raw output) was written); `run_backward` joins it.  Only with in-place arena gradients: a
        gradient tensor handed to autograd would have to cross streams."""
        if not (self.overlap_wgrad and dev.type == "cuda" and direct_grad(op.weight) is not None
                and (op.bias is None or direct_grad(op.bias) is not None)):
            yield
            return
        # MTB200_WGRAD_STREAMS side streams taken round-robin per layer: the deep-level launches are a few dozen CTAs each,
        # so more than two kernels can share the 148 SMs
        pool = self._side.get(dev)
        if pool is None:
            n = max(1, int(os.environ.get("MTB200_WGRAD_STREAMS", "2")))
            pool = self._side[dev] = [[torch.cuda.Stream(dev) for _ in range(n)], 0]
        side = pool[0][pool[1] % len(pool[0])]
        pool[1] += 1
        if after is not None:
            side.wait_event(after)
        else:
            side.wait_stream(torch.cuda.current_stream(dev))
        self._side_used = pool[0]
        with torch.cuda.stream(side):
            yield

    def _wgrad(self, tape, op: ConvOp, x: Feat, dy: Feat, dw, dt, dev, need_input_grad, bias_grad_is_zero):
        if self.use_c1(op) and x.xform is None and not need_input_grad:
            B, D, H, W = x.dims
            L.call("mtb200_conv_c1_wgrad", C.c_void_p(x.ptr() + x.coff * x.buf.element_size()), x.ldc, dy.ptr(), dy.ldc,
                   dy.coff, op.Cout_p, L.ptr(dw), op.Cin_p, dt, B, D, H, W, L.stream_ptr(),
                   flops=self.conv_flops(op, dy.dims), tag="conv_wgrad",
                   info=(1, op.Cout_p, tuple(dy.dims[1:]), op.ntap, (1, 1, 1), (1, 1, 1)))
            self._finish_wgrad(tape, op, dw, dy, dt, dev, bias_grad_is_zero)
            return
        if x.Cp != op.Cin_p:
            raise L.Mtb200Error("compact single-channel input: only the first-layer kernels can read it "
                                "(no input gradient, 16-bit tensor-core path)")
        p = L.WgradParams()
        p.x, p.dy, p.dw = x.ptr(), dy.ptr(), dw.data_ptr()
        p.xform = x.xform.data_ptr() if x.xform is not None else None
        p.dtype = dt
        p.B, p.Di, p.Hi, p.Wi = x.dims
        p.in_ldc, p.in_coff, p.Cin = x.ldc, x.coff, op.Cin_p
        p.in_split = x.split
        _, p.Dof, p.Hof, p.Wof = dy.dims
        p.out_ldc, p.out_coff, p.Cout = dy.ldc, dy.coff, op.Cout_p
        p.Do, p.Ho, p.Wo = x.dims[1:] if op.transposed else dy.dims[1:]
        if self._wpairs_convT(op, dy):
            # ConvTranspose3d(k == s, kx = 2) with 32 output channels: on the [.., W/2, 64] pair view of dy it is a
            # transposed convolution with kernel (k0, k1, 1) and 64 output channels (kx, co), and its weight gradient
            # [k0 k1][64][Cin] IS dw[(kz, ky, kx)][32][Cin] -- the same memory; the dY bricks become dense 128-byte rows
            p.Wof, p.out_ldc, p.Cout = dy.dims[3] // 2, 64, 64
            _pair_convT_taps(op.kernel).fill(p)
        else:
            op.fwd_taps.fill(p)
        p.impl = self.impl
        L.call("mtb200_wgrad_taps", C.byref(p), L.stream_ptr(), flops=self.conv_flops(op, dy.dims), tag="conv_wgrad",
               info=(op.Cin_p, op.Cout_p, (p.Do, p.Ho, p.Wo), op.ntap, op.fwd_taps.in_stride, op.fwd_taps.out_stride))
        self._finish_wgrad(tape, op, dw, dy, dt, dev, bias_grad_is_zero)

    def _conv_bwd(self, tape, op: ConvOp, x: Feat, dy: Feat, need_input_grad, bias_grad_is_zero=False):
        """Weight / bias gradient and data gradient of a (transposed) convolution given d(raw output)."""
        dev = dy.buf.device
        dt = L.dtype_enum(self.dtype)
        x = self.operand(x)
        # ---- weight gradient (same tap table as the forward problem)
        dw = self._z32.take((op.ntap, op.Cout_p, op.Cin_p), dev)
        # Launch order (MTB200_WGRAD_ORDER).  0: weight gradient first (as in round 1) -- its persistent CTAs take every SM
        # and the data gradient, which the rest of the backward chain waits for, queues behind it.  1 (default): data
        # gradient first; the weight gradient depends only on d(raw output) and is enqueued behind it on the side stream,
        # so it shares the SMs with the HBM-bound InstanceNorm-backward passes of the NEXT layer instead of delaying them.
        # 2: as 1, but the weight gradient also waits for the data gradient to finish.
        order = self.wgrad_order if (need_input_grad and self.overlap_wgrad and dev.type == "cuda") else 0
        ready = None
        if order == 1:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
        if order == 0:
            with self._wgrad_stream(op, dev):
                self._wgrad(tape, op, x, dy, dw, dt, dev, need_input_grad, bias_grad_is_zero)
        # ---- data gradient
        if not need_input_grad:
            return
        fl = self.conv_flops(op, dy.dims)
        gx, have = tape.grad_feat(x)
        grid = gx.dims[1:] if op.transposed else tuple(n // s for n, s in zip(gx.dims[1:], op.stride))
        dyv = Feat(dy.buf, dy.coff, dy.C, dy.Cp)  # gradients carry no pending transform
        # this launch writes the FINAL d(activation) of a single-consumer layer: offer the fused InstanceNorm-backward
        # reduction (sum dv, sum dv * xhat) to the kernel's epilogue
        red = None
        raw = x.raw_of() if x.raw_of is not None else None
        if (self.fuse_red and raw is not None and raw.single_consumer and not have and raw.xform is not None
                and raw.meanrstd is not None and raw.Cp == gx.Cp and raw.dims == gx.dims and self.materialize_inputs):
            red = (raw, self._z64.take((gx.dims[0], raw.Cp, 2), dev))
        if (self.wpairs_dgrad and self.wpairs and not op.transposed and op.kernel == (3, 3, 3) and op.stride[2] == 2
                and op.Cin_p == 32 and gx.ldc == 32 and gx.coff == 0 and gx.buf.is_contiguous() and gx.dims[3] % 2 == 0
                and (gx.planar is None or gx.half is not None) and op.Cout_p in (16, 32, 64)
                and self.dtype in (torch.bfloat16, torch.float16) and self.impl in (0, 2)):
            # data gradient of the strided 3x3x3 convolution into a dense 32-channel tensor: on the pair view of d_in the 8
            # residue groups become 4 groups of 64 channels (rw, ci) -- 128-byte rows for the group-merged kernel's store
            B, D, H, W = gx.dims
            gxv = Feat(gx.buf.view(B, D, H, W // 2, 64), 0, 64, 64)
            wd = op.packed(self.wdtype, True)
            ver = (op.weight._version, _weights_epoch, op.weight.data_ptr(), wd.data_ptr())
            hit = op._packed.get("wpairD")
            if hit is None or hit[0] != ver:
                hit = (ver, wpair_weights_conv_dgrad(wd, hit[1] if hit is not None else None))
                op._packed["wpairD"] = hit
            took = self._conv_call(wpair_taps_conv_dgrad(op.stride), dyv, hit[1], None, gxv, grid, None, have, op.Cout_p,
                                   64, flops=fl, tag="conv_dgrad", red=None)
        elif self._wpairs_convT(op, dy):
            # data gradient of ConvTranspose3d(k == s) with 32 output channels: d_in[q] = sum_k W[k]^T dy[s q + k]; the
            # taps kx = 0, 1 are ONE 128-byte row of the [.., W/2, 64] pair view of dy -- k0 k1 taps of K = 64 on dense
            # rows instead of k0 k1 2 taps of K = 32 on every other 64-byte row
            B, D, H, W = dyv.dims
            dyp = Feat(dyv.buf.view(B, D, H, W // 2, 64), 0, 64, 64)
            took = self._conv_call(wpair_taps_convT_dgrad(op.kernel), dyp, self._wpair_weights_convT(op), None, gx, grid,
                                   None, have, 64, op.Cin_p, flops=fl, tag="conv_dgrad", red=red)
        else:
            took = self._conv_call(op.dgrad_taps, dyv, op.packed(self.wdtype, True), None, gx, grid, None, have,
                                   op.Cout_p, op.Cin_p, flops=fl, tag="conv_dgrad", red=red)
        if took:
            raw.red_fused = red[1]
        tape.mark(x)
        if order:
            with self._wgrad_stream(op, dev, after=ready):
                self._wgrad(tape, op, x, dy, dw, dt, dev, need_input_grad, bias_grad_is_zero)

    def _head_bwd_fused(self, tape, op: ConvOp, x: Feat, y: Feat, spec):
        """d(loss)/d(logits) (never stored) -> head data gradient + head weight gradient, one launch."""
        dev = y.buf.device
        dt = L.dtype_enum(self.dtype)
        x = self.operand(x)
        gx, have = tape.grad_feat(x)
        dw = self._z32.take((op.ntap, op.Cout_p, op.Cin_p), dev)
        p = L.HeadBwdParams()
        p.logits = None if spec.get("deferred") else y.ptr()  # deferred: the kernel recomputes the window from x
        p.w_fwd = op.packed(self.wdtype, False).data_ptr()
        p.target, p.coef = spec["target"].data_ptr(), spec["coef"].data_ptr()
        p.gscale, p.pos_mask = spec["gscale"].data_ptr(), spec["pos"].data_ptr()
        p.x, p.w_swap, p.dx, p.dw = x.ptr(), op.packed(self.wdtype, True).data_ptr(), gx.ptr(), dw.data_ptr()
        p.nvox, p.dtype, p.B = y.nvox, dt, y.dims[0]
        p.z_ldc, p.C8, p.n_labels = y.ldc, spec["C8"], spec["n_labels"]
        p.x_ldc, p.x_coff, p.Cin, p.Cout = x.ldc, x.coff, op.Cin_p, op.Cout_p
        p.dx_ldc, p.dx_coff, p.accumulate = gx.ldc, gx.coff, int(have)
        for b, c0 in enumerate(spec["win"]):
            p.win_c0[b] = c0
        L.call("mtb200_head_bwd_fused", C.byref(p), L.stream_ptr(), flops=2 * self.conv_flops(op, y.dims),
               tag="conv_head_bwd", info=(op.Cin_p, op.Cout_p, tuple(y.dims[1:]), 1, (1, 1, 1), (1, 1, 1)))
        tape.mark(x)
        tape.keep.append(spec)
        self._finish_wgrad(tape, op, dw, None, dt, dev, bias_grad_is_zero=True)

    def _finish_wgrad(self, tape, op: ConvOp, dw, dy: Feat, dt, dev, bias_grad_is_zero):
        """Packed fp32 weight gradient -> the parameter's gradient (arena slot or autograd), plus the bias gradient."""
        gw = direct_grad(op.weight)
        if gw is not None and self.batch_unpack and gw.is_contiguous():
            self._pending_unpack.append((op, dw, gw))  # folded in by run_backward's single batched launch
            tape.direct_done.add(id(op.weight))
        elif gw is not None:  # accumulate straight into the arena's gradient slot
            L.call("mtb200_unpack_wgrad", L.ptr(dw), op.Cout, op.Cin, op.ntap, int(op.transposed), op.Cout_p, op.Cin_p,
                   op.split, op.split_p, 1.0, 1, L.ptr(gw), L.stream_ptr())
            tape.direct_done.add(id(op.weight))
        else:
            gw = torch.empty_like(op.weight)
            L.call("mtb200_unpack_wgrad", L.ptr(dw), op.Cout, op.Cin, op.ntap, int(op.transposed), op.Cout_p, op.Cin_p,
                   op.split, op.split_p, 1.0, 0, L.ptr(gw), L.stream_ptr())
            tape.add_param_grad(op.weight, gw)
        if op.bias is not None and bias_grad_is_zero:
            if direct_grad(op.bias) is not None:
                tape.direct_done.add(id(op.bias))
            else:
                tape.add_param_grad(op.bias, torch.zeros_like(op.bias))
        elif op.bias is not None:
            gb = torch.zeros(op.Cout_p, dtype=torch.float32, device=dev)
            L.call("mtb200_colsum", dy.ptr(), dt, dy.dims[0] * dy.nvox, dy.ldc, dy.coff, op.Cout_p, L.ptr(gb),
                   L.stream_ptr())
            tape.add_param_grad(op.bias, gb[:op.Cout])

    def conv_plain(self, tape: Optional[Tape], op: ConvOp, x: Feat, out: Optional[Feat] = None,
                   need_input_grad=True, head=False) -> Feat:
        """Convolution with no normalisation after it (transposed-conv upsampling, 1x1x1 heads).  `head`: a segmentation
        head whose output goes to the loss -- its backward may arrive as a `lazy_heads` descriptor (fused kernel)."""
        fusable = (head and tape is not None and self.fuse_head and need_input_grad and op.bias is None and op.ntap == 1
                   and not op.transposed and op.Cin_p in (32, 64) and self.materialize_inputs and out is None
                   and x.dims[0] <= L.MAX_HEAD_BATCH)
        if (head and tape is None and self.capture_head and op.ntap == 1 and not op.transposed and op.Cin_p in (32, 64)
                and op.Cout_p <= 48 and self.materialize_inputs and out is None):
            self.captured_head = (self.operand(x), op)  # the predictor consumes the head's input itself
            ph = torch.zeros(8, dtype=self.dtype, device=x.buf.device)[:1].expand(tuple(x.dims) + (op.Cout_p,))
            return Feat(ph, 0, op.Cout, op.Cout_p)
        deferred = fusable and self.defer_heads
        if deferred:
            # not computed: a zero-stride placeholder of the logits' shape stands in (its one-element storage is the key)
            ph = torch.zeros(8, dtype=self.dtype, device=x.buf.device)[:1].expand(tuple(x.dims) + (op.Cout_p,))
            y = Feat(ph, 0, op.Cout, op.Cout_p)
            self.deferred[ph.data_ptr()] = {"x": x, "op": op}
        else:
            y, _ = self.conv(op, x, out)
        if fusable:
            self.fusable_heads[y.buf.data_ptr()] = True
        if tape is not None:
            def bwd():
                spec = self.lazy_heads.pop(y.buf.data_ptr(), None)
                if spec is not None:
                    self._head_bwd_fused(tape, op, x, y, spec)
                    return
                if deferred:
                    raise L.Mtb200Error("a deferred head's output reached the backward pass without the fused loss "
                                        "(multitalent_loss(..., engine=network._engine) must consume it)")
                if not tape.has_grad(y):
                    self._zero_param_grads(tape, op)
                    return
                g, _ = tape.grad_feat(y)
                self._conv_bwd(tape, op, x, g, need_input_grad)
            tape.closures.append(bwd)
        return y

    def residual_act(self, tape: Optional[Tape], a: Feat, r: Feat, slope2=LRELU_SLOPE, out: Optional[Feat] = None) -> Feat:
        """out = LeakyReLU(f(a) + g(r)) -- tail of BasicResidualBlock.forward (conv_blocks.py:205-213).  `out` may be a
        channel slice of a wider buffer (the skip half of a decoder concat buffer)."""
        out = self.materialize(a, res=r, slope2=slope2, out=out)
        if tape is not None:
            def bwd():
                if not tape.has_grad(out):
                    return
                g, _ = tape.grad_feat(out)
                ga, have_a = tape.grad_feat(a)
                gr, have_r = tape.grad_feat(r)
                B = out.dims[0]
                L.call("mtb200_residual_bwd", g.ptr(), g.ldc, g.coff, out.ptr(), out.ldc, out.coff,
                       ga.ptr(), ga.ldc, ga.coff, int(have_a), gr.ptr(), gr.ldc, gr.coff, int(have_r),
                       L.dtype_enum(self.dtype), B * out.nvox, out.Cp, float(slope2), L.stream_ptr())
                tape.mark(a)
                tape.mark(r)
            tape.closures.append(bwd)
        return out

    def seed_grad(self, tape: Tape, y: Feat, dy: torch.Tensor):
        """Install d(loss)/d(y) for an output feature.  `dy` is NCDHW-shaped; a channels-last view of an NDHWC buffer of
        the engine dtype is adopted without a copy."""
        ldc = ndhwc_view_info(dy) if dy.dtype == self.dtype else None
        if ldc is not None and ldc == y.ldc and y.coff == 0:
            B, Cc, D, H, W = dy.shape
            buf = torch.as_strided(dy, (B, D, H, W, ldc), (D * H * W * ldc, H * W * ldc, W * ldc, ldc, 1))
            tape.grad_bufs[id(y.buf)] = buf
            tape.grad_init[id(y.buf)] = set()
            tape.keep.append(dy)
        else:
            g, _ = tape.grad_feat(y)
            B, Cc, D, H, W = dy.shape
            src = dy.detach().float().contiguous()
            L.call("mtb200_ncdhw_to_ndhwc", L.ptr(src), B, Cc, D * H * W, g.ptr(), L.dtype_enum(self.dtype), g.ldc,
                   g.coff, g.Cp, L.stream_ptr())
            tape.keep.append(src)
        tape.mark(y)

    def run_backward(self, tape: Tape):
        if self.bwd_priority and torch.cuda.is_available():
            # the dgrad / InstanceNorm-backward chain on a high-priority stream: when an SM frees up, its blocks are
            # placed before the queued weight-gradient CTAs of the side streams
            cur = torch.cuda.current_stream()
            hp = self._hp.get(cur.device)
            if hp is None:
                hp = self._hp[cur.device] = torch.cuda.Stream(cur.device, priority=-1)
            hp.wait_stream(cur)
            with torch.cuda.stream(hp):
                for c in reversed(tape.closures):
                    c()
            cur.wait_stream(hp)
        else:
            for c in reversed(tape.closures):
                c()
        tape.closures = []
        if self._side_used is not None:  # the optimizer (and the next forward's pool reset) must see every weight gradient
            for st in self._side_used:
                torch.cuda.current_stream().wait_stream(st)
            self._side_used = None
        self._flush_unpack()

    def side_streams_in_use(self):
        return list(self._side_used) if self._side_used is not None else []

    def _flush_unpack(self):
        """grad += dw for every layer of this step in one launch.  The descriptor table lives on the device and is cached
        by the (dw, grad) pointers: the pools hand out the same addresses every step, so there is no per-step H2D copy."""
        pend, self._pending_unpack = self._pending_unpack, []
        if not pend:
            return
        key = tuple((dw.data_ptr(), gw.data_ptr()) for _, dw, gw in pend)
        hit = self._unpack_tables.get(key)
        if hit is None:
            if len(self._unpack_tables) > 8:
                self._unpack_tables.clear()
            descs = (L.UnpackDesc * len(pend))()
            blk = 0
            for d, (op, dw, gw) in zip(descs, pend):
                d.dw, d.grad = dw.data_ptr(), gw.data_ptr()
                d.Cout, d.Cin, d.ntap, d.transposed = op.Cout, op.Cin, op.ntap, int(op.transposed)
                d.Cout_p, d.Cin_p, d.split, d.split_p = op.Cout_p, op.Cin_p, op.split, op.split_p
                d.blk_begin = blk
                blk += (op.Cout * op.Cin * op.ntap + L.UNPACK_CHUNK - 1) // L.UNPACK_CHUNK
            table = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(pend[0][1].device)
            hit = self._unpack_tables[key] = (table, len(pend), blk)
        table, n, blocks = hit
        L.call("mtb200_unpack_wgrad_batched", L.ptr(table), n, blocks, L.stream_ptr(), tag="mtb200_unpack_wgrad")
