"""Fold ensembling of the sliding-window predictor -- the inner block of `predict_cases`
(nnunet/inference/predict_MultiTalent.py:222-240): for every fold load its parameters, predict the preprocessed case,
sum the per-class probability volumes, divide by the number of folds and undo the plan's axis transposition.

The reference round-trips every fold's 47 x X x Y x Z float32 volume through host memory (25 GB for 512^3) and adds them
with numpy; here the running sum stays in HBM (one fp32 volume) and a single D2H copy is made at the end.  Everything
around this block in the reference (file lists, preprocessing workers, NIfTI export) is out of scope.
"""
from typing import Sequence

import numpy as np
import torch


def predict_case_with_folds(trainer, params: Sequence[dict], d: np.ndarray, do_tta: bool = True, step_size: float = 0.5,
                            all_in_gpu: bool = False, mixed_precision: bool = True, mirror_axes=(0, 1, 2),
                            return_device_tensor: bool = False):
    """`params` = one checkpoint dict per fold (what `load_model_and_checkpoint_files` returns, model_restore.py:149);
    `d` = the preprocessed case (c, x, y, z).  Returns the fold-averaged probabilities [47, x, y, z] (numpy float32, or
    a CUDA tensor with `return_device_tensor`) in the ORIGINAL axis order (transpose_backward applied, :237-240)."""
    assert len(params) >= 1, "need at least one fold"
    net = trainer.network
    total = None
    for p in params:
        trainer.load_checkpoint_ram(p, False)
        ds, mode = net.do_ds, net.training
        net.do_ds = False
        net.eval()
        try:
            _, prob = net.predict_3D(d, do_mirroring=do_tta, mirror_axes=tuple(mirror_axes) if do_tta else (),
                                     use_sliding_window=True, step_size=step_size, patch_size=tuple(trainer.patch_size),
                                     regions_class_order=trainer.regions_class_order, use_gaussian=True,
                                     all_in_gpu=all_in_gpu, verbose=False, mixed_precision=mixed_precision,
                                     return_device_tensors=True)
        finally:
            net.train(mode)
            net.do_ds = ds
        if total is None:
            total = prob
        else:
            total += prob
    if len(params) > 1:
        total /= float(len(params))
    tf = trainer.plans.get('transpose_forward') if trainer.plans is not None else None
    if tf is not None:
        tb = trainer.plans.get('transpose_backward')
        total = total.permute([0] + [i + 1 for i in tb])
    if return_device_tensor:
        return total
    return total.contiguous().cpu().numpy()
