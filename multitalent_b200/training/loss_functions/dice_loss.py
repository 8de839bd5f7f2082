"""Generic nnU-Net segmentation loss on the native kernels: `DC_and_CE_loss` (softmax + cross entropy + soft Dice) --
replaces nnunet/training/loss_functions/dice_loss.py:488-545 (`DC_and_CE_loss`), :155-195 (`SoftDiceLoss`), :100-152
(`get_tp_fp_fn_tn`), nd_softmax.py:20 (`softmax_helper`) and crossentropy.py:4-11 (`RobustCrossEntropyLoss`) for the
configuration every nnUNetTrainerV2-style trainer uses ({'batch_dice': ..., 'smooth': 1e-5, 'do_bg': False}, {},
aggregate "sum", weights 1, no squared Dice, no ignore label; nnUNetTrainer.py:134).  Used by the fine-tuning trainers
of the MultiTalent repository, not by the MultiTalent trainers themselves.

    loss = CE(z, y) - mean_{c in classes'} (2 tp_c + s) / (2 tp_c + fp_c + fn_c + s + 1e-8)

Native form: pass 1 = per-(sample, class) {tp, sum p, count} + CE sum (one kernel), a few [B, C]-sized tensor ops for
the Dice value and its d/dp coefficients, pass 2 = d(loss)/d(logits) (one kernel).  With `group` the statistics are
summed over the ranks first (the pooling of nnUNetTrainerV2_DDP.compute_loss, nnUNetTrainerV2_DDP.py:249-282, through
`awesome_allgather_function`): the backward then carries the factor world_size, see multitalent_loss.py.
"""
import torch
import torch.distributed as dist
from torch import nn

from ... import _lib as L
from ...engine import pad_channels
from .multitalent_loss import _as_ndhwc, pool_stats_over_ranks


class _DcCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, batch_dice, do_bg, smooth, eps, weight_ce, weight_dice, group):
        dev = logits.device
        B, Cc = logits.shape[:2]
        nvox = logits.shape[2] * logits.shape[3] * logits.shape[4]
        dt = logits.dtype if logits.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
        zv, ldc = _as_ndhwc(logits, dt)
        Cp = (Cc + 7) // 8 * 8
        if Cp > 64:
            raise L.Mtb200Error("DC_and_CE_loss: at most 64 classes on the native path (got %d)" % Cc)
        tgt = target.detach().float().contiguous()
        assert tgt.numel() == B * nvox, "target %s does not match logits %s" % (tuple(target.shape), tuple(logits.shape))
        stats = torch.zeros((B, Cp, 3), dtype=torch.float64, device=dev)
        ce_sum = torch.zeros(B, dtype=torch.float64, device=dev)
        st = L.stream_ptr()
        L.call("mtb200_dcce_stats", L.ptr(zv), L.dtype_enum(dt), ldc, Cc, Cp, L.ptr(tgt), B, nvox, L.ptr(stats),
               L.ptr(ce_sum), st)
        world = dist.get_world_size(group) if (group is not None and dist.is_available() and dist.is_initialized()) else 1
        pooled = pool_stats_over_ranks(stats, group) if world > 1 else stats
        s = pooled[:, :Cc]                                     # [B, C, 3] = tp, sum p, count
        if batch_dice:
            s = s.sum(0, keepdim=True)
        tp, sp, cnt = s[..., 0], s[..., 1], s[..., 2]
        nom = 2 * tp + smooth
        den = sp + cnt + smooth + eps                           # 2tp + fp + fn = sum p + count
        dc = nom / den
        c0 = 0 if do_bg else 1
        n_terms = dc[:, c0:].numel()
        dice_loss = -dc[:, c0:].sum() / n_terms
        ce = ce_sum.sum() / float(B * nvox)
        loss = weight_ce * ce + weight_dice * dice_loss
        # d(dice_loss)/dp_c at a voxel = a_c [y == c] + b_c
        a = (-2.0 / den) * (weight_dice / n_terms) * world
        bb = (nom / (den * den)) * (weight_dice / n_terms) * world
        coef = torch.zeros((B, Cp, 2), dtype=torch.float32, device=dev)
        coef[:, c0:Cc, 0] = a[:, c0:].float().expand(B, -1)
        coef[:, c0:Cc, 1] = bb[:, c0:].float().expand(B, -1)
        ctx.save = (zv, ldc, dt, tgt, coef, B, Cc, Cp, nvox, tuple(logits.shape), float(weight_ce) / float(B * nvox))
        return loss.float()

    @staticmethod
    def backward(ctx, g):
        zv, ldc, dt, tgt, coef, B, Cc, Cp, nvox, shape, ce_w = ctx.save
        _, _, D, H, W = shape
        dz = torch.empty((B, D, H, W, ldc), dtype=dt, device=zv.device)
        if ldc > Cp:
            dz[..., Cp:].zero_()
        gs = g.detach().float().reshape(1).contiguous()
        L.call("mtb200_dcce_bwd", L.ptr(zv), L.dtype_enum(dt), ldc, Cc, Cp, L.ptr(tgt), B, nvox, L.ptr(coef), ce_w,
               L.ptr(gs), L.ptr(dz), ldc, L.stream_ptr())
        return (dz.permute(0, 4, 1, 2, 3)[:, :Cc],) + (None,) * 8


class DC_and_CE_loss(nn.Module):
    """Same constructor as the reference (dice_loss.py:489-516).  Settings outside the configuration the nnU-Net
    trainers use raise: there is no silent eager fallback on the native path."""

    def __init__(self, soft_dice_kwargs, ce_kwargs, aggregate="sum", square_dice=False, weight_ce=1, weight_dice=1,
                 log_dice=False, ignore_label=None, group=None):
        super().__init__()
        if aggregate != "sum" or square_dice or log_dice or ignore_label is not None or ce_kwargs:
            raise NotImplementedError("DC_and_CE_loss: only aggregate='sum', plain soft Dice, no ignore label, default CE")
        kw = dict(soft_dice_kwargs)
        self.batch_dice = bool(kw.pop('batch_dice', False))
        self.do_bg = bool(kw.pop('do_bg', True))
        self.smooth = float(kw.pop('smooth', 1.))
        if kw:
            raise NotImplementedError("DC_and_CE_loss: unsupported soft_dice_kwargs %s" % sorted(kw))
        self.weight_ce, self.weight_dice, self.group = weight_ce, weight_dice, group

    def forward(self, net_output, target):
        if not net_output.is_cuda:
            raise L.Mtb200Error("DC_and_CE_loss runs on the native CUDA path only; logits are on %s" % net_output.device)
        return _DcCeFn.apply(net_output, target, self.batch_dice, self.do_bg, self.smooth, 1e-8, self.weight_ce,
                             self.weight_dice, self.group)
