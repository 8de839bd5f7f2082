"""Deep-supervision target down-sampling on the device -- `downsample_seg_for_ds_transform2` /
`DownsampleSegForDSTransform2` of nnunet/training/data_augmentation/downsampling.py:70-104.

The reference resizes every (b, c) label volume on the CPU inside the augmentation workers with batchgenerators'
`resize_segmentation(seg, new_shape, order=0)` = skimage.transform.resize(order 0, mode 'edge', anti_aliasing False):
nearest neighbour with pixel-centre coordinates, src = floor((o + 0.5) * in / out).  Here the label batch is already in
HBM (it is an input of the loss) and each scale is one launch of `mtb200_resize_nearest`."""
import numpy as np
import torch

from ... import _lib as L


def downsample_seg_for_ds_transform2(seg, ds_scales=((1, 1, 1), (0.5, 0.5, 0.5), (0.25, 0.25, 0.25)), order=0, axes=None):
    """`seg`: CUDA float32 tensor [B, C, X, Y, Z] (label maps).  Returns the list of resized tensors, one per scale; a scale
    of all ones returns `seg` itself (downsampling.py:93-94).  Only `order=0` (what nnU-Net uses for targets,
    data_augmentation_moreDA.py:193-198) is implemented."""
    if order != 0:
        raise NotImplementedError("deep-supervision targets are resized with order 0 (nearest neighbour)")
    if not seg.is_cuda:
        raise L.Mtb200Error("downsample_seg_for_ds_transform2 runs on the native CUDA path only; got %s" % seg.device)
    if axes is None:
        axes = list(range(2, seg.dim()))
    assert seg.dim() == 5 and list(axes) == [2, 3, 4], "3D label batches [B, C, X, Y, Z]"
    src = seg.detach().float().contiguous()
    out = []
    for s in ds_scales:
        if all(i == 1 for i in s):
            out.append(seg)
            continue
        new_shape = np.array(src.shape).astype(float)
        for i, a in enumerate(axes):
            new_shape[a] *= s[i]
        new_shape = np.round(new_shape).astype(int)                       # downsampling.py:96-99
        dst = torch.empty(tuple(int(v) for v in new_shape), dtype=torch.float32, device=src.device)
        L.call("mtb200_resize_nearest", L.ptr(src), int(src.shape[0] * src.shape[1]), int(src.shape[2]), int(src.shape[3]),
               int(src.shape[4]), L.ptr(dst), int(new_shape[2]), int(new_shape[3]), int(new_shape[4]), L.stream_ptr())
        out.append(dst)
    return out


class DownsampleSegForDSTransform2(object):
    """downsampling.py:70-84: `data_dict[output_key]` becomes the list of targets at `ds_scales`."""

    def __init__(self, ds_scales=(1, 0.5, 0.25), order=0, input_key="seg", output_key="seg", axes=None):
        self.axes, self.output_key, self.input_key, self.order, self.ds_scales = axes, output_key, input_key, order, ds_scales

    def __call__(self, **data_dict):
        data_dict[self.output_key] = downsample_seg_for_ds_transform2(data_dict[self.input_key], self.ds_scales, self.order,
                                                                      self.axes)
        return data_dict
