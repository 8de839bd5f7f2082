"""Online (per-iteration) hard-Dice evaluation of the MultiTalent trainers --
`MultiTalent_trainer_ddp.run_online_evaluation` / `finish_online_evaluation`
(nnunet/training/network_training/custom_trainers/MultiTalent/MultiTalent/MultiTalent_Trainer_DDP.py:372-430).

Per sample b and supervised region r (channel j_r): prediction = sigmoid(z) > 0.5 (i.e. z > 0), ground truth =
OR_{l in regions[r]} (target == l); tp / fp / fn counted over the voxels of the highest-resolution output; channels the
sample's dataset does not label stay 0.  The reference walks a Python double loop with one masked sum per (b, r).  Here,
inside a trainer step, the three counts come out of the loss's own statistics kernel (`mtb200_mt_loss_stats`, `hard`
output: the validation iteration reads the logits once for loss AND evaluation); called stand-alone, the region membership
of every voxel is ONE table lookup (`label -> 47 booleans`) and the counts are boolean reductions in plain torch ops on
whatever device the tensors live on.  The counts are exact integers, so the result equals the reference's float32 sums
below 2^24 voxels per (b, region)."""
from typing import Sequence

import numpy as np
import torch
import torch.distributed as dist

from ..dataset_conversion.Task100_MultiTalent import (NUM_LABELS, NUM_OUTPUT_CHANNELS, MultiTalent_region_output_idx_mapping,
                                                      MultiTalent_regions)

_TABLE = {}


def _membership_table(device):
    t = _TABLE.get(device)
    if t is None:
        m = np.zeros((NUM_LABELS, NUM_OUTPUT_CHANNELS), dtype=bool)
        for name, labels in MultiTalent_regions.items():
            m[list(labels), MultiTalent_region_output_idx_mapping[name]] = True
        t = _TABLE[device] = torch.from_numpy(m).to(device)
    return t


def hard_tp_fp_fn(output0: torch.Tensor, target0: torch.Tensor, valid_regions: Sequence[Sequence[str]]):
    """`output0` [B, 47, D, H, W] logits (any strides / float dtype), `target0` [B, 1, D, H, W] float label map ->
    three float32 tensors [B, 47] (MT:381-397)."""
    B, C = output0.shape[:2]
    dev = output0.device
    table = _membership_table(dev)
    tp = torch.zeros((B, C), dtype=torch.float32, device=dev)
    fp = torch.zeros_like(tp)
    fn = torch.zeros_like(tp)
    with torch.no_grad():
        for b in range(B):
            chans = sorted({MultiTalent_region_output_idx_mapping[r] for r in valid_regions[b]})
            if not chans:
                continue
            idx = torch.tensor(chans, device=dev)
            lab = target0[b, 0].long().clamp_(0, NUM_LABELS - 1).reshape(-1)
            pred = (output0[b].index_select(0, idx) > 0).reshape(len(chans), -1)       # [n, voxels]
            gt = table.index_select(1, idx)[lab].t()                                   # [n, voxels]
            tp[b, idx] = (pred & gt).sum(1).float()
            fp[b, idx] = (pred & ~gt).sum(1).float()
            fn[b, idx] = (~pred & gt).sum(1).float()
    return tp, fp, fn


def _gather_ranks(t: torch.Tensor, group=None) -> torch.Tensor:
    """Forward of awesome_allgather_function (utilities/distributed.py:28-47): [W, *t.shape]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t[None]
    bufs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(bufs, t.contiguous(), group=group)
    return torch.stack(bufs, 0)


class OnlineEvaluationMixin:
    """The four accumulator lists and the two methods of the reference trainer, same names and contents."""

    def _online_eval_reset(self):
        self.online_eval_foreground_dc, self.online_eval_tp = [], []
        self.online_eval_fp, self.online_eval_fn = [], []

    def run_online_evaluation(self, output, target, valid_regions):
        """MT:372-410."""
        if not hasattr(self, "online_eval_tp"):
            self._online_eval_reset()
            self.all_val_eval_metrics = getattr(self, "all_val_eval_metrics", [])
        fused = getattr(self, "_hard_stats", None)
        if fused and fused.get("logits") is not None and fused["logits"].data_ptr() == output[0].data_ptr():
            tp, fp, fn = fused["tp"], fused["fp"], fused["fn"]   # counted by the loss kernel of this very step
        else:
            tp, fp, fn = hard_tp_fp_fn(output[0], target[0], valid_regions)
        self._hard_stats = None
        tp_hard = _gather_ranks(tp).cpu().numpy()
        fp_hard = _gather_ranks(fp).cpu().numpy()
        fn_hard = _gather_ranks(fn).cpu().numpy()
        self.online_eval_foreground_dc.append(list((2 * tp_hard) / (2 * tp_hard + fp_hard + fn_hard + 1e-8)))
        self.online_eval_tp.append(list(tp_hard.sum(0)))
        self.online_eval_fp.append(list(fp_hard.sum(0)))
        self.online_eval_fn.append(list(fn_hard.sum(0)))

    def finish_online_evaluation(self):
        """MT:412-430: global Dice per channel over everything accumulated since the last call; appends the mean to
        `all_val_eval_metrics` and returns the per-channel list."""
        tp = np.sum(self.online_eval_tp, 0)
        fp = np.sum(self.online_eval_fp, 0)
        fn = np.sum(self.online_eval_fn, 0)
        per_class = [l for l in [2 * i / (np.clip(2 * i + j + k, a_min=1e-8, a_max=None)) for i, j, k in zip(tp, fp, fn)]
                     if not np.isnan(l).any()]
        self.all_val_eval_metrics.append(np.mean(per_class))
        self._online_eval_reset()
        return per_class
