"""CPU tests: the C-ABI library loads and exports exactly what include/mtb200.h declares; host logic (tap tables,
module tree / state_dict keys, trainer bookkeeping, no-fallback behaviour)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
from torch import nn

from conftest import ROOT, build_small_net

from multitalent_b200 import _lib as L
from multitalent_b200 import engine as E


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mtb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        from multitalent_b200.build import build
        build()
    lib = ctypes.CDLL(L.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), "libmtb200.so does not export %s" % n
    assert sorted(L.SIGNATURES) == names, "ctypes binding and header disagree"
    lib.mtb200_version.restype = ctypes.c_int
    assert lib.mtb200_version() == 100


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of the parameter blocks as the C compiler lays them out from include/mtb200.h == the ctypes
    Structures of multitalent_b200/_lib.py (guards against ctypes/C drift)."""
    import subprocess
    fields = {"mtb200_conv_params": (L.ConvParams, ["in", "stats", "dtype", "Cin", "Do", "ngroups", "ntaps", "tap_widx",
                                                     "accumulate", "red_y", "red", "red_ldc", "impl", "in_split", "out_split"]),
              "mtb200_wgrad_params": (L.WgradParams, ["x", "xform", "dtype", "Cout", "ntaps", "tap_widx", "impl", "in_split"]),
              "mtb200_head_bwd_params": (L.HeadBwdParams, ["logits", "w_fwd", "dw", "nvox", "dtype", "Cin", "accumulate",
                                                           "win_c0"]),
              "mtb200_head_agg_params": (L.HeadAggParams, ["x", "nb", "weight", "dtype", "flip", "z0"]),
              "mtb200_head_fwd_params": (L.HeadFwdParams, ["x", "hard", "nvox", "dtype", "Cout", "win_c0"]),
              "mtb200_pack_desc": (L.PackDesc, ["w", "packed_swap", "Cout", "blk_begin"]),
              "mtb200_unpack_desc": (L.UnpackDesc, ["dw", "grad", "Cout", "blk_begin"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mtb200.h"', 'int main(void) {']
    for st, (_, fl) in fields.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (st, st))
        for f in fl:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (st, f, st, f))
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    got = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    rename = {"in": "inp", "is": "is_", "os": "os_"}
    for st, (cls, fl) in fields.items():
        assert int(got[st]) == ctypes.sizeof(cls), st
        for f in fl:
            assert int(got["%s.%s" % (st, f)]) == getattr(cls, rename.get(f, f)).offset, (st, f)


def test_pad_channels():
    assert [E.pad_channels(c) for c in (1, 30, 47, 60, 120, 240, 320)] == [16, 32, 48, 64, 128, 256, 320]


def _dense_conv_from_taps(table, x, w, out_shape, grid):
    """Evaluate a tap table with numpy loops (tiny sizes) -- checks the tables against torch's conv semantics."""
    B, Cin = x.shape[0], x.shape[1]
    Cout = w.shape[1]
    out = np.zeros((B, Cout) + tuple(out_shape))
    for g, ooff in enumerate(table.group_ooff):
        for t in range(table.group_begin[g], table.group_begin[g + 1]):
            off, widx = table.taps[t]
            for d in range(grid[0]):
                for h in range(grid[1]):
                    for ww in range(grid[2]):
                        src = [d * table.in_stride[0] + off[0], h * table.in_stride[1] + off[1],
                               ww * table.in_stride[2] + off[2]]
                        if any(s < 0 or s >= n for s, n in zip(src, x.shape[2:])):
                            continue
                        dst = (d * table.out_stride[0] + ooff[0], h * table.out_stride[1] + ooff[1],
                               ww * table.out_stride[2] + ooff[2])
                        out[(slice(None), slice(None)) + dst] += x[:, :, src[0], src[1], src[2]] @ w[widx].T
    return out


@pytest.mark.parametrize("kernel,stride", [((3, 3, 3), (1, 1, 1)), ((3, 3, 3), (2, 2, 2)), ((3, 3, 3), (1, 2, 2)),
                                           ((1, 3, 3), (1, 1, 1)), ((1, 1, 1), (2, 2, 2))])
def test_tap_tables_conv(kernel, stride):
    torch.manual_seed(0)
    x = torch.randn(1, 3, 4, 6, 4, dtype=torch.float64, requires_grad=True)
    w = torch.randn(2, 3, *kernel, dtype=torch.float64)
    pad = [(k - 1) // 2 for k in kernel]
    y = torch.nn.functional.conv3d(x, w, stride=stride, padding=pad)
    wp = w.permute(2, 3, 4, 0, 1).reshape(-1, 2, 3).numpy()  # [tap][co][ci]
    got = _dense_conv_from_taps(E.taps_conv_fwd(kernel, stride), x.detach().numpy(), wp, y.shape[2:], y.shape[2:])
    np.testing.assert_allclose(got, y.detach().numpy(), atol=1e-10)
    gy = torch.randn_like(y)
    y.backward(gy)
    wpt = np.transpose(wp, (0, 2, 1))  # [tap][ci][co]
    grid = [n // s for n, s in zip(x.shape[2:], stride)]
    got = _dense_conv_from_taps(E.taps_conv_dgrad(kernel, stride), gy.numpy(), wpt, x.shape[2:], grid)
    np.testing.assert_allclose(got, x.grad.numpy(), atol=1e-10)


def _wpair_view(t):
    """NCDHW [B, C, D, H, W] -> the w-pair view [B, 2C, D, H, W/2] with channel index pw * C + c (what the NDHWC buffer
    looks like when two neighbouring voxels are read as one row)."""
    B, Cc, D, H, W = t.shape
    return np.ascontiguousarray(t.reshape(B, Cc, D, H, W // 2, 2).transpose(0, 5, 1, 2, 3, 4).reshape(B, 2 * Cc, D, H, W // 2))


@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2)])
def test_wpair_tables_strided_conv(stride):
    """Engine.conv's stride-2 forward on the pair view: 18 taps of 2C channels == Conv3d(3x3x3, stride (.., 2))."""
    torch.manual_seed(1)
    x = torch.randn(1, 3, 4, 6, 8, dtype=torch.float64)
    w = torch.randn(2, 3, 3, 3, 3, dtype=torch.float64)
    y = torch.nn.functional.conv3d(x, w, stride=stride, padding=1)
    wp = w.permute(2, 3, 4, 0, 1).reshape(27, 2, 3).contiguous()  # [tap][co][ci]
    wv = E.wpair_weights(wp).reshape(18, 2, 6).numpy()
    got = _dense_conv_from_taps(E.wpair_taps(stride), _wpair_view(x.numpy()), wv, y.shape[2:], y.shape[2:])
    np.testing.assert_allclose(got, y.numpy(), atol=1e-10)


@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2)])
def test_wpair_tables_strided_conv_dgrad(stride):
    """Data gradient of the strided convolution written through the pair view of d_in (4 groups of 2C channels)."""
    torch.manual_seed(3)
    x = torch.randn(1, 3, 4, 6, 8, dtype=torch.float64, requires_grad=True)
    w = torch.randn(2, 3, 3, 3, 3, dtype=torch.float64)
    y = torch.nn.functional.conv3d(x, w, stride=stride, padding=1)
    gy = torch.randn_like(y)
    y.backward(gy)
    wd = w.permute(2, 3, 4, 1, 0).reshape(27, 3, 2).contiguous()  # [tap][ci][co]
    wv = E.wpair_weights_conv_dgrad(wd).reshape(18, 6, 2).numpy()
    gv_shape = (x.shape[2], x.shape[3], x.shape[4] // 2)
    got = _dense_conv_from_taps(E.wpair_taps_conv_dgrad(stride), gy.numpy(), wv, gv_shape, y.shape[2:])
    np.testing.assert_allclose(got, _wpair_view(x.grad.numpy()), atol=1e-10)


@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2)])
def test_wpair_dgrad_table_fits_the_group_merged_kernel(stride):
    """csrc/conv_gm.cu's envelope, restated: the groups are exactly the residue classes of the output lattice, no group
    has two taps at one input offset, at most 32 distinct offsets -- otherwise the pair-view data gradient would fall
    back to the per-tap kernel without anyone noticing."""
    t = E.wpair_taps_conv_dgrad(stride)
    assert t.in_stride == (1, 1, 1) and t.out_stride == (stride[0], stride[1], 1)
    assert sorted(t.group_ooff) == sorted((a, b, 0) for a in range(stride[0]) for b in range(stride[1]))
    offsets = set()
    for g in range(len(t.group_ooff)):
        offs = [t.taps[i][0] for i in range(t.group_begin[g], t.group_begin[g + 1])]
        assert len(offs) >= 1 and len(set(offs)) == len(offs)
        offsets.update(offs)
    assert len(offsets) <= 32 and len(t.taps) == 18
    assert sorted(w for _, w in t.taps) == list(range(18))  # every virtual weight slice used exactly once


@pytest.mark.parametrize("kernel", [(2, 2, 2), (1, 2, 2)])
def test_wpair_tables_conv_transpose(kernel):
    """ConvTranspose3d(k == s, kx = 2) on the pair view of its output (forward) and of dy (data gradient); the forward
    weights and the weight gradient need no repacking: [(kz, ky, kx)][co][ci] is [(kz, ky)][(kx, co)][ci] as it lies."""
    torch.manual_seed(2)
    x = torch.randn(1, 3, 2, 3, 2, dtype=torch.float64, requires_grad=True)
    w = torch.randn(3, 2, *kernel, dtype=torch.float64)  # [Cin][Cout][k]
    y = torch.nn.functional.conv_transpose3d(x, w, stride=kernel)
    wp = w.permute(2, 3, 4, 1, 0).reshape(-1, 2, 3).contiguous()  # [tap][co][ci]
    table = E._pair_convT_taps(kernel)
    yv_shape = (y.shape[2], y.shape[3], y.shape[4] // 2)
    got = _dense_conv_from_taps(table, x.detach().numpy(), wp.reshape(-1, 4, 3).numpy(), yv_shape, x.shape[2:])
    np.testing.assert_allclose(got, _wpair_view(y.detach().numpy()), atol=1e-10)
    gy = torch.randn_like(y)
    y.backward(gy)
    wd = wp.permute(0, 2, 1).contiguous()  # [tap][ci][co]
    wv = E.wpair_weights_convT_dgrad(wd, kernel).reshape(-1, 3, 4).numpy()
    got = _dense_conv_from_taps(E.wpair_taps_convT_dgrad(kernel), _wpair_view(gy.numpy()), wv, x.shape[2:], x.shape[2:])
    np.testing.assert_allclose(got, x.grad.numpy(), atol=1e-10)


@pytest.mark.parametrize("kernel", [(2, 2, 2), (1, 2, 2)])
def test_tap_tables_conv_transpose(kernel):
    torch.manual_seed(0)
    x = torch.randn(1, 3, 2, 3, 2, dtype=torch.float64, requires_grad=True)
    w = torch.randn(3, 2, *kernel, dtype=torch.float64)  # [Cin][Cout][k]
    y = torch.nn.functional.conv_transpose3d(x, w, stride=kernel)
    wp = w.permute(2, 3, 4, 1, 0).reshape(-1, 2, 3).numpy()  # [tap][co][ci]
    got = _dense_conv_from_taps(E.taps_convT_fwd(kernel), x.detach().numpy(), wp, y.shape[2:], x.shape[2:])
    np.testing.assert_allclose(got, y.detach().numpy(), atol=1e-10)
    gy = torch.randn_like(y)
    y.backward(gy)
    got = _dense_conv_from_taps(E.taps_convT_dgrad(kernel), gy.numpy(), np.transpose(wp, (0, 2, 1)), x.shape[2:],
                                x.shape[2:])
    np.testing.assert_allclose(got, x.grad.numpy(), atol=1e-10)


def test_module_tree_matches_fixture_keys(golden_small):
    blob, meta = golden_small
    net = build_small_net(meta, blob, device="cpu")
    keys = [k[len("param/"):] for k in blob if k.startswith("param/")]
    assert list(net.state_dict().keys()) == keys  # same names AND same registration order as the reference
    assert net._native_ok and net.do_ds and net._deep_supervision
    assert list(net.input_shape_must_be_divisible_by) == [4, 8, 8]
    assert net.conv_op == nn.Conv3d and net.num_classes == 47


def test_no_cpu_fallback_on_native_configuration(golden_small):
    blob, meta = golden_small
    net = build_small_net(meta, blob, device="cpu")
    with pytest.raises(L.Mtb200Error):
        net(torch.zeros(1, 1, 8, 16, 16))
    from multitalent_b200.training.loss_functions.multitalent_loss import multitalent_loss
    with pytest.raises(L.Mtb200Error):
        multitalent_loss([torch.zeros(1, 47, 8, 16, 16)], [torch.zeros(1, 1, 8, 16, 16)], [("03_liver",)], [1.0])


def test_non_native_configuration_uses_torch_modules():
    """2D / BatchNorm / dropout configurations are outside the native path and run through the torch leaf modules."""
    from multitalent_b200.network_architecture.generic_UNet import Generic_UNet
    net = Generic_UNet(1, 4, 3, 2, conv_op=nn.Conv2d, norm_op=nn.BatchNorm2d, dropout_op=nn.Dropout2d,
                       final_nonlin=lambda x: x, convolutional_pooling=False, convolutional_upsampling=False)
    assert not net._native_ok
    out = net(torch.zeros(2, 1, 16, 16))
    assert out[0].shape == (2, 3, 16, 16)


def test_trainer_bookkeeping():
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import (
        MultiTalent_trainer_ddp, nnUNetTrainerV2_MultiTalent, poly_lr)
    assert nnUNetTrainerV2_MultiTalent is MultiTalent_trainer_ddp
    t = MultiTalent_trainer_ddp(None, 0, 0, init_distributed=False)
    t.initialize(True)
    np.testing.assert_allclose(t.ds_loss_weights, np.array([8, 4, 2, 1, 0]) / 15)
    assert [list(map(float, s)) for s in t.deep_supervision_scales] == [[1, 1, 1], [.5] * 3, [.25] * 3, [.125] * 3,
                                                                         [1 / 16] * 3]
    assert t.num_classes == 47 and len(t.init_args) == 11
    assert abs(poly_lr(500, 1000, 1e-2) - 1e-2 * 0.5 ** 0.9) < 1e-12
    n_params = sum(p.numel() for p in t.network.parameters())
    assert n_params == 29319560
    assert len(t.network.state_dict()) == 98


def test_region_bitmasks():
    from multitalent_b200.dataset_conversion.Task100_MultiTalent import region_bitmasks, valid_channel_mask
    pos, chan = region_bitmasks()
    assert len(pos) == 48 and pos[0] == 0
    assert pos[2] == (1 << chan['03_liver']) | (1 << chan['03_cancer'])  # label 2 (liver tumour) is in both regions
    assert pos[43] == (1 << chan['64_both_kidneys']) | (1 << chan['64_kidney_tumor'])
    assert valid_channel_mask(('03_liver', '03_cancer')) == 0b11
    assert bin(valid_channel_mask(tuple(chan))).count("1") == 47


def test_head_windows_cover_every_dataset():
    """Host logic of the fused head kernels: every Task100 dataset's supervised channels fit one 16-channel window that
    starts on a multiple of 8 and ends inside the padded 48 channels; a spread-out (synthetic) region set does not."""
    from multitalent_b200.dataset_conversion.Task100_MultiTalent import (MultiTalent_region_output_idx_mapping,
                                                                          MultiTalent_task_ids, MultiTalent_valid_regions,
                                                                          valid_channel_mask)
    from multitalent_b200.training.loss_functions.multitalent_loss import head_windows
    regs = [MultiTalent_valid_regions[t] for t in MultiTalent_task_ids]
    win = head_windows(regs, 48)
    assert win is not None and len(win) == len(regs)
    for c0, r in zip(win, regs):
        m = valid_channel_mask(r)
        assert c0 % 8 == 0 and 0 <= c0 and c0 + 16 <= 48
        assert m >> (c0 + 16) == 0 and m & ((1 << c0) - 1) == 0, (c0, bin(m))
    names = sorted(MultiTalent_region_output_idx_mapping, key=MultiTalent_region_output_idx_mapping.get)
    assert head_windows([(names[0], names[-1])], 48) is None     # channels 0 and 46: no 16-channel window
    assert head_windows([()], 48) == (0,)                        # nothing supervised: any window
