"""Export-side resampling of the predicted probabilities on the device -- the arithmetic of
`save_segmentation_nifti_from_softmax` (nnunet/inference/segmentation_export.py:27-159) up to the file writing, and of
`resample_data_or_seg(is_seg=False)` (nnunet/preprocessing/preprocessing.py:109-197) which it calls (SURVEY.md section
8(f) N2).

The reference does this per region in a process pool on the CPU (47 single-channel volumes per case,
predict_MultiTalent.py:252-263): skimage.transform.resize(order 1, mode 'edge', anti_aliasing False) -- in-plane only
plus order-0 interpolation along the low-resolution axis when the spacing is anisotropic ("separate z") -- then the
in-order threshold `seg[p_i > 0.5] = c_i` and the paste into the uncropped volume.  Here one `mtb200_resample_probs`
launch reads the probability volume where the predictor left it (HBM), interpolates with per-axis order 0 / 1 (pixel-centre
coordinates, edge clamping), writes the label map and, if asked, the resampled probabilities as fp16 (what the
reference's `.npz` holds).  NIfTI writing stays in the reference (SimpleITK)."""
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L

RESAMPLING_SEPARATE_Z_ANISO_THRESHOLD = 3   # nnunet/configuration.py:4


def get_do_separate_z(spacing, anisotropy_threshold=RESAMPLING_SEPARATE_Z_ANISO_THRESHOLD):
    """preprocessing.py:30-32."""
    return (np.max(spacing) / np.min(spacing)) > anisotropy_threshold


def get_lowres_axis(new_spacing):
    """preprocessing.py:35-37."""
    return np.where(max(new_spacing) / np.array(new_spacing) == 1)[0]


def resample_softmax_and_threshold(segmentation_softmax: torch.Tensor, properties_dict: dict, order: int = 1,
                                   region_class_order: Optional[Sequence] = None, force_separate_z: Optional[bool] = None,
                                   interpolation_order_z: int = 0, return_probabilities: bool = False):
    """segmentation_export.py:77-139 on a CUDA fp32 tensor [C, x, y, z].  Returns (seg uint8 at
    `original_size_of_raw_data` with the crop bbox pasted, resampled probabilities fp16 [C, X, Y, Z] or None)."""
    if not segmentation_softmax.is_cuda:
        raise L.Mtb200Error("resample_softmax_and_threshold runs on the native CUDA path only; got %s" %
                            segmentation_softmax.device)
    if order not in (0, 1) or interpolation_order_z not in (0, 1):
        raise NotImplementedError("interpolation orders 0 and 1 (the MultiTalent export default is 1 / 0)")
    src = segmentation_softmax.detach().float().contiguous()
    C = int(src.shape[0])
    cur = tuple(int(v) for v in src.shape[1:])
    after_crop = tuple(int(v) for v in properties_dict.get('size_after_cropping'))
    before_crop = properties_dict.get('original_size_of_raw_data')
    orders = [order] * 3
    if any(i != j for i, j in zip(cur, after_crop)):
        if force_separate_z is None:
            if get_do_separate_z(properties_dict.get('original_spacing')):
                do_sep, axis = True, get_lowres_axis(properties_dict.get('original_spacing'))
            elif get_do_separate_z(properties_dict.get('spacing_after_resampling')):
                do_sep, axis = True, get_lowres_axis(properties_dict.get('spacing_after_resampling'))
            else:
                do_sep, axis = False, None
        else:
            do_sep = force_separate_z
            axis = get_lowres_axis(properties_dict.get('original_spacing')) if do_sep else None
        if axis is not None and len(axis) != 1:
            do_sep = False
        if do_sep:
            orders[int(axis[0])] = interpolation_order_z     # in-plane `order`, `order_z` along the low-res axis
    dev = src.device
    n = after_crop
    seg = torch.empty(n, dtype=torch.uint8, device=dev)
    prob = torch.empty((C,) + n, dtype=torch.float16, device=dev) if return_probabilities else None
    co = None
    if region_class_order is not None:
        co = torch.tensor([float(c) for c in region_class_order], dtype=torch.float32, device=dev)
        assert co.numel() == C
    L.call("mtb200_resample_probs", L.ptr(src), C, cur[0], cur[1], cur[2], n[0], n[1], n[2], orders[0], orders[1],
           orders[2], L.ptr(prob), 1, L.ptr(co), L.ptr(seg), L.stream_ptr())
    bbox = properties_dict.get('crop_bbox')
    if bbox is not None:                                      # :125-133
        full = torch.zeros(tuple(int(v) for v in before_crop), dtype=torch.uint8, device=dev)
        b = [[int(bb[0]), int(min(bb[0] + seg.shape[c], before_crop[c]))] for c, bb in enumerate(bbox)]
        full[b[0][0]:b[0][1], b[1][0]:b[1][1], b[2][0]:b[2][1]] = seg[:b[0][1] - b[0][0], :b[1][1] - b[1][0], :b[2][1] - b[2][0]]
        seg = full
    return seg, prob
