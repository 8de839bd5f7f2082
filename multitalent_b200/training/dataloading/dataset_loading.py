"""Patch sampling of `DataLoader3D.generate_train_batch` (nnunet/training/dataloading/dataset_loading.py:224-380) with the
crop + pad on the device.

Host side = the reference's statement sequence, restated so that it consumes `np.random` in the same order: case choice
with the 1/sqrt(n_dataset) sampling probabilities (MultiTalent_Trainer_DDP.py:629-633, 651-657), foreground oversampling
of the last `oversample_foreground_percent` of the batch, bounding-box bounds from `need_to_pad`, class / voxel choice from
the pre-computed `class_locations`.  Device side: one `mtb200_crop_pad` launch per sample takes the patch out of the
case volume that already lives in HBM -- data channels padded with `pad_kwargs_data` (or the edge voxel), the label
channel with -1 -- instead of the slice copy + two `np.pad` calls per sample in the CPU worker processes."""
from collections import OrderedDict

import numpy as np
import torch

from ... import _lib as L


def crop_and_pad_case(case_all_data: torch.Tensor, bbox_lb, patch_size, pad_mode="constant", pad_kwargs_data=None,
                      out_data=None, out_seg=None):
    """`case_all_data`: CUDA float32 [c + 1, X, Y, Z] (last channel = label map).  Returns (data [c, *patch], seg [1,
    *patch]) = dataset_loading.py:353-368: np.pad of the valid part with `pad_mode` / `pad_kwargs_data` for the data
    channels and constant -1 for the label channel."""
    if not case_all_data.is_cuda:
        raise L.Mtb200Error("crop_and_pad_case runs on the native CUDA path only; got %s" % case_all_data.device)
    if pad_mode not in ("constant", "edge"):
        raise NotImplementedError("pad_mode %r (the MultiTalent loaders use 'constant', DataLoader3D defaults to 'edge')" % pad_mode)
    src = case_all_data.detach().float().contiguous()
    C, X, Y, Z = (int(v) for v in src.shape)
    pd, ph, pw = (int(v) for v in patch_size)
    dev = src.device
    data = out_data if out_data is not None else torch.empty((C - 1, pd, ph, pw), dtype=torch.float32, device=dev)
    seg = out_seg if out_seg is not None else torch.empty((1, pd, ph, pw), dtype=torch.float32, device=dev)
    cv = float((pad_kwargs_data or {}).get('constant_values', 0))
    st = L.stream_ptr()
    lb = [int(v) for v in bbox_lb]
    if C > 1:
        pv = torch.full((C - 1,), cv, dtype=torch.float32, device=dev)
        L.call("mtb200_crop_pad", L.ptr(src), C - 1, X, Y, Z, lb[0], lb[1], lb[2], L.ptr(data), pd, ph, pw,
               int(pad_mode == "edge"), L.ptr(pv), st)
    minus1 = torch.full((1,), -1.0, dtype=torch.float32, device=dev)
    L.call("mtb200_crop_pad", L.ptr(src[C - 1:]), 1, X, Y, Z, lb[0], lb[1], lb[2], L.ptr(seg), pd, ph, pw, 0,
           L.ptr(minus1), st)
    return data, seg


class DataLoader3D(object):
    """The 3D patch loader over cases that are resident in HBM.  `data`: {key: {'data': CUDA tensor [c + 1, X, Y, Z],
    'properties': dict with 'class_locations'}} (the reference reads `data_file` npy/npz from disk).  Same constructor
    arguments and the same `generate_train_batch()` result dictionary as the reference class (:154-200, :224, :380) with
    CUDA tensors for 'data' and 'seg'."""

    def __init__(self, data, patch_size, final_patch_size, batch_size, has_prev_stage=False,
                 oversample_foreground_percent=0.0, memmap_mode="r", pad_mode="edge", pad_kwargs_data=None,
                 pad_sides=None, sampling_probabilities=None):
        if has_prev_stage:
            raise NotImplementedError("cascade (seg_from_prev_stage) is not on the MultiTalent path")
        self._data, self.batch_size = data, batch_size
        self.pad_kwargs_data = pad_kwargs_data if pad_kwargs_data is not None else OrderedDict()
        self.pad_mode = pad_mode
        self.oversample_foreground_percent = oversample_foreground_percent
        self.final_patch_size, self.patch_size = final_patch_size, patch_size
        self.list_of_keys = list(self._data.keys())
        self.need_to_pad = (np.array(patch_size) - np.array(final_patch_size)).astype(int)
        if pad_sides is not None:
            self.need_to_pad += np.array(pad_sides)
        self.sampling_probabilities = sampling_probabilities
        k = self.list_of_keys[0]
        self.data_shape = (batch_size, int(self._data[k]['data'].shape[0]) - 1, *patch_size)
        self.seg_shape = (batch_size, 1, *patch_size)

    _crop = staticmethod(crop_and_pad_case)   # the device kernel (tests swap in the oracle's numpy restatement on the CPU)

    def get_do_oversample(self, batch_idx):
        return not batch_idx < round(self.batch_size * (1 - self.oversample_foreground_percent))

    def sample_bbox(self, shape, properties, force_fg):
        """dataset_loading.py:271-335: lower corner of the patch (may be negative / overhang: the rest is padding)."""
        need_to_pad = self.need_to_pad.copy()
        for d in range(3):
            if need_to_pad[d] + shape[d] < self.patch_size[d]:
                need_to_pad[d] = self.patch_size[d] - shape[d]
        lb = [-need_to_pad[d] // 2 for d in range(3)]
        ub = [shape[d] + need_to_pad[d] // 2 + need_to_pad[d] % 2 - self.patch_size[d] for d in range(3)]
        if not force_fg:
            return [np.random.randint(lb[d], ub[d] + 1) for d in range(3)]
        if 'class_locations' not in properties.keys():
            raise RuntimeError("Please rerun the preprocessing with the newest version of nnU-Net!")
        fg = np.array([i for i in properties['class_locations'].keys() if len(properties['class_locations'][i]) != 0])
        fg = fg[fg > 0]
        if len(fg) == 0:
            return [np.random.randint(lb[d], ub[d] + 1) for d in range(3)]
        selected_class = np.random.choice(fg)
        voxels = properties['class_locations'][selected_class]
        v = voxels[np.random.choice(len(voxels))]
        return [max(lb[d], int(v[d]) - self.patch_size[d] // 2) for d in range(3)]

    def generate_train_batch(self):
        selected_keys = np.random.choice(self.list_of_keys, self.batch_size, True, self.sampling_probabilities)
        dev = self._data[self.list_of_keys[0]]['data'].device
        data = torch.empty(self.data_shape, dtype=torch.float32, device=dev)
        seg = torch.empty(self.seg_shape, dtype=torch.float32, device=dev)
        case_properties = []
        for j, i in enumerate(selected_keys):
            force_fg = self.get_do_oversample(j)
            properties = self._data[i]['properties']
            case_properties.append(properties)
            case_all_data = self._data[i]['data']
            bbox = self.sample_bbox(tuple(int(v) for v in case_all_data.shape[1:]), properties, force_fg)
            self._crop(case_all_data, bbox, self.patch_size, self.pad_mode, self.pad_kwargs_data,
                       out_data=data[j], out_seg=seg[j])
        return {'data': data, 'seg': seg, 'properties': case_properties, 'keys': selected_keys}

    def __next__(self):
        return self.generate_train_batch()

    def __iter__(self):
        return self
