"""MultiTalent multi-head loss on the native kernels -- replaces `MultiTalent_trainer_ddp.compute_loss`
(nnunet/training/network_training/custom_trainers/MultiTalent/MultiTalent/MultiTalent_Trainer_DDP.py:544-623) and the
differentiable all-gather it uses (nnunet/utilities/distributed.py:28-73).

Per deep-supervision scale i (weight w_i, MT:85-96):  CE_i = sum over (b, r in valid_regions[b]) of the voxel-mean
BCE-with-logits of channel j_r against y = OR_{l in regions[r]} (target == l);  tp/fp/fn[b, j_r] = sums of sigma*y,
sigma*(1-y), (1-sigma)*y;  all-gathered over ranks and summed over the RANK axis only (pooling per local batch index b,
MT:596-604);  DC_i = sum_{b,j} 2tp / clamp(2tp + fp + fn, 1e-7);  loss = sum_i w_i (CE_i - DC_i).

Native form: two streaming kernels per scale.  Pass 1 accumulates {sum bce, tp, sum sigma, sum y} per (b, j)
(note 2tp + fp + fn = sum sigma + sum y).  ONE all-gather of the packed [scales, B, C, 2] buffer replaces the
reference's 15 tiny collectives; the backward needs NO collective: every rank holds the same pooled Dice, so the
all-reduced gradient (distributed.py:71) is exactly world_size x the local closed form, which pass 2 evaluates:
    dL/dz = w_i [ (sigma - y)/N_vox  -  W * sigma (1 - sigma) (2 y D - 2 TP) / D^2 ]        for supervised (b, j), else 0.
"""
import ctypes
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from ... import _lib as L
from ...dataset_conversion.Task100_MultiTalent import NUM_LABELS, region_bitmasks, valid_channel_mask
from ...engine import ndhwc_view_info, pad_channels

_POS_MASK_CACHE = {}


def _pos_mask(device):
    t = _POS_MASK_CACHE.get(device)
    if t is None:
        pos, _ = region_bitmasks()
        t = torch.tensor(pos, dtype=torch.int64, device=device)
        _POS_MASK_CACHE[device] = t
    return t


def pool_stats_over_ranks(packed: torch.Tensor, group=None) -> torch.Tensor:
    """`packed` [..., 2] = this rank's {tp, sum sigma + sum y}; returns the sum over ranks (all_gather + sum(0), the
    forward of awesome_allgather_function followed by `.sum(0, keepdim=True)`, MT:598-604).  Works on any backend
    (NCCL on GPUs, gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return packed
    bufs = [torch.empty_like(packed) for _ in range(dist.get_world_size(group))]
    dist.all_gather(bufs, packed.contiguous(), group=group)
    return torch.stack(bufs, 0).sum(0)


def _as_ndhwc(t: torch.Tensor, dtype):
    """(buffer-like tensor, ldc) for an NCDHW-shaped logits tensor; adopts channels-last views, converts otherwise."""
    if t.dtype == dtype:
        ldc = ndhwc_view_info(t)
        if ldc is not None:
            return t, ldc
    B, Cc, D, H, W = t.shape
    Cp = pad_channels(Cc)
    src = t.detach().float().contiguous()
    buf = torch.empty((B, D, H, W, Cp), dtype=dtype, device=t.device)
    L.call("mtb200_ncdhw_to_ndhwc", L.ptr(src), B, Cc, D * H * W, L.ptr(buf), L.dtype_enum(dtype), Cp, 0, Cp,
           L.stream_ptr())
    return buf.permute(0, 4, 1, 2, 3)[:, :Cc], Cp


class _MultiTalentLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, valid_mask, weights, group, targets, hard_out, lazy, *logits):
        dev = logits[0].device
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        pos = _pos_mask(dev)
        st = L.stream_ptr()
        B, Cc = logits[0].shape[:2]
        C8 = (Cc + 7) // 8 * 8
        active = [i for i, w in enumerate(weights) if w != 0]
        views, stats = {}, {}
        for i in active:
            z = logits[i]
            dt = z.dtype if z.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
            dspec = lazy[0].deferred.get(z.data_ptr()) if lazy is not None else None
            if dspec is not None and lazy[1] is None:
                raise L.Mtb200Error("deferred head, but the supervised channels of this batch do not fit the 16-channel "
                                    "windows of the fused kernels")
            if dspec is not None:
                # deferred head: no logits exist.  Pass 1 runs on the head's input (tensor core), mtb200_head_fwd_stats.
                eng, win = lazy
                tgt = targets[i].detach().float().contiguous()
                nvox = z.shape[2] * z.shape[3] * z.shape[4]
                assert tgt.shape[0] == B and tgt.numel() == B * nvox, "target does not match the deferred logits"
                s = torch.zeros((B, C8, 4), dtype=torch.float64, device=dev)
                hard = None
                if hard_out is not None and i == 0:
                    hard = torch.zeros((B, C8, 2), dtype=torch.float64, device=dev)
                x, op = eng.operand(dspec["x"]), dspec["op"]
                assert op.Cout_p == C8, (op.Cout_p, C8)
                p = L.HeadFwdParams()
                p.x, p.w_fwd, p.target = x.ptr(), op.packed(eng.wdtype, False).data_ptr(), tgt.data_ptr()
                p.valid_mask, p.pos_mask, p.stats = valid_mask.data_ptr(), pos.data_ptr(), s.data_ptr()
                p.hard = hard.data_ptr() if hard is not None else None
                p.nvox, p.dtype, p.B, p.C8, p.n_labels = nvox, L.dtype_enum(dt), B, C8, NUM_LABELS
                p.x_ldc, p.x_coff, p.Cin, p.Cout = x.ldc, x.coff, op.Cin_p, op.Cout_p
                for b, c0 in enumerate(win):
                    p.win_c0[b] = c0
                L.call("mtb200_head_fwd_stats", ctypes.byref(p), st, flops=eng.conv_flops(op, (B,) + tuple(z.shape[2:])),
                       tag="conv_head_fwd", info=(op.Cin_p, op.Cout_p, tuple(z.shape[2:]), 1, (1, 1, 1), (1, 1, 1)))
                if hard is not None:
                    tp = hard[:, :Cc, 0]
                    hard_out.update(tp=tp.float(), fp=(hard[:, :Cc, 1] - tp).float(), fn=(s[:, :Cc, 3] - tp).float(),
                                    logits=None)
                views[i] = (z, C8, dt, tgt, nvox)
                stats[i] = s
                continue
            zv, ldc = _as_ndhwc(z, dt)
            tgt = targets[i]
            assert tgt.shape[0] == B and tgt.numel() == B * z.shape[2] * z.shape[3] * z.shape[4], \
                "target %s does not match logits %s" % (tuple(tgt.shape), tuple(z.shape))
            tgt = tgt.detach().float().contiguous()
            nvox = z.shape[2] * z.shape[3] * z.shape[4]
            s = torch.zeros((B, C8, 4), dtype=torch.float64, device=dev)
            hard = None
            if hard_out is not None and i == 0:  # online evaluation reads the highest-resolution output (MT:373-374)
                hard = torch.zeros((B, C8, 2), dtype=torch.float64, device=dev)
            L.call("mtb200_mt_loss_stats", L.ptr(zv), L.dtype_enum(dt), ldc, C8, L.ptr(tgt), B, nvox, L.ptr(valid_mask),
                   L.ptr(pos), NUM_LABELS, L.ptr(s), L.ptr(hard), st)
            if hard is not None:
                tp = hard[:, :Cc, 0]
                hard_out.update(tp=tp.float(), fp=(hard[:, :Cc, 1] - tp).float(), fn=(s[:, :Cc, 3] - tp).float(),
                                logits=logits[0])
            views[i] = (zv, ldc, dt, tgt, nvox)
            stats[i] = s
        pooled = {i: None for i in active}
        if world > 1 and active:
            packed = torch.stack([torch.stack((stats[i][..., 1], stats[i][..., 2] + stats[i][..., 3]), -1)
                                  for i in active], 0)
            allp = pool_stats_over_ranks(packed, group)
            pooled = {i: allp[k].contiguous() for k, i in enumerate(active)}
        losses = torch.zeros(3, dtype=torch.float32, device=dev)
        coefs = {}
        for i in active:
            coef = torch.empty((B, C8, 4), dtype=torch.float32, device=dev)
            L.call("mtb200_mt_loss_finalize", L.ptr(stats[i]), L.ptr(pooled[i]), L.ptr(valid_mask), B, C8, views[i][4],
                   float(weights[i]), float(world), L.ptr(losses), L.ptr(coef), st)
            coefs[i] = coef
        ctx.views, ctx.coefs, ctx.active, ctx.n = views, coefs, active, len(logits)
        ctx.shapes = [tuple(z.shape) for z in logits]
        ctx.pos = pos
        ctx.C8 = C8
        ctx.lazy = lazy
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_loss, g_ce, g_dc):
        st = L.stream_ptr()
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n
        gs = g_loss.detach().float().contiguous()
        for i in ctx.active:
            zv, ldc, dt, tgt, nvox = ctx.views[i]
            B, Cc, D, H, W = ctx.shapes[i]
            if ctx.lazy is not None and ctx.lazy[1] is not None:
                # the network's engine runs this head's backward as ONE kernel (loss pass 2 + head data gradient + head
                # weight gradient, mtb200_head_bwd_fused): d(logits) is never written.  What flows through autograd is a
                # zero-stride placeholder of the right shape; the descriptor travels beside it, keyed by the buffer.
                eng, win = ctx.lazy
                deferred = zv.data_ptr() in eng.deferred
                if (eng.fusable_heads.get(zv.data_ptr()) and dt in (torch.bfloat16, torch.float16) and ldc == ctx.C8
                        and ctx.needs_input_grad[6 + i]):
                    eng.lazy_heads[zv.data_ptr()] = {"target": tgt, "coef": ctx.coefs[i], "gscale": gs, "pos": ctx.pos,
                                                     "C8": ctx.C8, "n_labels": NUM_LABELS, "win": win,
                                                     "deferred": deferred}
                    grads[i] = torch.zeros((), dtype=dt, device=zv.device).expand(B, Cc, D, H, W)
                    continue
            if ctx.lazy is not None and zv.data_ptr() in ctx.lazy[0].deferred:
                raise L.Mtb200Error("deferred head: the fused backward is not available for this output")
            dz = torch.empty((B, D, H, W, ldc), dtype=dt, device=zv.device)
            if ldc > ctx.C8:
                dz[..., ctx.C8:].zero_()
            L.call("mtb200_mt_loss_bwd", L.ptr(zv), L.dtype_enum(dt), ldc, ctx.C8, L.ptr(tgt), B, nvox, L.ptr(ctx.pos),
                   NUM_LABELS, L.ptr(ctx.coefs[i]), L.ptr(gs), L.ptr(dz), ldc, st)
            grads[i] = dz.permute(0, 4, 1, 2, 3)[:, :Cc]
        for i in range(ctx.n):
            if grads[i] is None and ctx.needs_input_grad[6 + i]:
                B, Cc, D, H, W = ctx.shapes[i]
                Cp = pad_channels(Cc)
                grads[i] = torch.zeros((B, D, H, W, Cp), dtype=ctx.views[ctx.active[0]][2],
                                       device=g_loss.device).permute(0, 4, 1, 2, 3)[:, :Cc]
        return (None, None, None, None, None, None) + tuple(grads)


_VALID_MASK_CACHE = {}


def valid_mask_tensor(valid_regions: Sequence[Sequence[str]], device) -> torch.Tensor:
    """[B] int64 bitmasks of supervised channels from the per-sample region-name tuples
    (`data_dict['properties'][b]['valid_regions']`, MT:328-329).  Cached per combination of datasets in the batch: a
    fresh `torch.tensor(..., device=cuda)` is a pageable H2D copy that blocks the host until the whole forward pass has
    drained, which would stop the launch queue from running ahead of the GPU once per step."""
    key = (str(device), tuple(tuple(v) for v in valid_regions))
    t = _VALID_MASK_CACHE.get(key)
    if t is None:
        if len(_VALID_MASK_CACHE) > 4096:
            _VALID_MASK_CACHE.clear()
        t = torch.tensor([valid_channel_mask(v) for v in valid_regions], dtype=torch.int64, device=device)
        _VALID_MASK_CACHE[key] = t
    return t


_WINDOW_CACHE = {}


def head_windows(valid_regions, C8, width=16):
    """Per sample: first channel (a multiple of 8) of a `width`-channel window that holds every supervised output channel
    of the sample's dataset (the regions of one dataset are consecutive output channels, Task100 tables), or None when
    some sample's channels do not fit one window -- then the heads run the unfused passes."""
    key = (tuple(tuple(v) for v in valid_regions), C8, width)
    if key not in _WINDOW_CACHE:
        if len(_WINDOW_CACHE) > 4096:
            _WINDOW_CACHE.clear()
        win = []
        for v in valid_regions:
            m = valid_channel_mask(v)
            if m == 0:
                win.append(0)
                continue
            first, last = (m & -m).bit_length() - 1, m.bit_length() - 1
            c0 = max(0, min(first // 8 * 8, C8 - width))
            if last >= c0 + width or first < c0:
                win = None
                break
            win.append(c0)
        _WINDOW_CACHE[key] = tuple(win) if win is not None else None
    return _WINDOW_CACHE[key]


def multitalent_loss(outputs, targets, valid_regions, ds_loss_weights, group=None, hard_out=None, engine=None):
    """(total_loss, total_ce, total_dc) as 0-dim CUDA tensors; `total_loss` is differentiable w.r.t. `outputs`.
    Only d(total_loss) is propagated (the trainer calls `l.backward()`; ce/dc are reporting values, MT:370).
    `engine`: the `Engine` of the network that produced `outputs` (optional) -- enables the fused head backward.
    `hard_out` (a dict, optional): filled with the hard tp / fp / fn [B, 47] of the highest-resolution output (the
    counts `run_online_evaluation` needs, MT:372-397) computed inside the loss's own statistics pass."""
    if not isinstance(outputs, (tuple, list)):
        outputs, targets = (outputs,), (targets,) if not isinstance(targets, (tuple, list)) else targets
    if not outputs[0].is_cuda:
        raise L.Mtb200Error("multitalent_loss runs on the native CUDA path only; logits are on %s" % outputs[0].device)
    vm = valid_regions if torch.is_tensor(valid_regions) else valid_mask_tensor(valid_regions, outputs[0].device)
    weights = [float(w) for w in ds_loss_weights][:len(outputs)]
    # `engine` (the network's kernel layer): lets the heads' backward run fused with the loss's second pass
    lazy = None
    if engine is not None and getattr(engine, "fuse_head", False) and not torch.is_tensor(valid_regions):
        lazy = (engine, head_windows(valid_regions, (outputs[0].shape[1] + 7) // 8 * 8))
    elif engine is not None and getattr(engine, "deferred", None):
        lazy = (engine, None)  # deferred heads present but no window information: the forward pass below refuses
    return _MultiTalentLossFn.apply(vm, weights, group, list(targets), hard_out, lazy, *outputs)
