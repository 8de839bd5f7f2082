"""`validate` of the MultiTalent trainers -- the compute half of
nnunet/training/network_training/custom_trainers/MultiTalent/MultiTalent/MultiTalent_Trainer_DDP.py:129-322.

What the reference does per validation case: rank-sharded case list (`all_keys[local_rank::world]`, :200), sliding-window
prediction (:231-235), per-dataset selection of the valid output channels and assembly of one label map by in-order
thresholding with the dataset's `regions_class_order` (:279-286 -> segmentation_export.py:118-123), file export through a
process pool (NIfTI, resampled to the original spacing), then `aggregate_scores` over the written files on rank 0
(:305-315).  Here: the same case sharding, prediction, channel selection and label-map assembly, with the probabilities
staying on the device; Dice per label is counted on the device against the case's ground-truth label map (the last
channel of the preprocessed `data` array, :226-229) and the per-case results are gathered to every rank.  File export
(NIfTI writing, resampling to the original spacing -- SURVEY.md section 8(f) N2) is delegated to `export_fn`; without it
nothing is written and the scores refer to the preprocessed grid rather than the original one.
"""
import pickle
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

from ..dataset_conversion.Task100_MultiTalent import (MultiTalent_region_output_idx_mapping, MultiTalent_regions_class_order,
                                                      MultiTalent_valid_regions)


def dataset_name_of_case(key: str) -> str:
    """MT:206-209: the case identifier starts with the numeric id of its source dataset."""
    names = [i for i in MultiTalent_valid_regions.keys() if i.startswith("Task%03.0d_" % int(key.split('_')[0]))]
    assert len(names) == 1, "cannot map case %r to one MultiTalent dataset" % key
    return names[0]


def assemble_label_map(probs: torch.Tensor, regions_class_order) -> torch.Tensor:
    """segmentation_export.py:118-123: seg = 0; for i, c in enumerate(regions_class_order): seg[probs[i] > 0.5] = c."""
    seg = torch.zeros(probs.shape[1:], dtype=torch.float32, device=probs.device)
    for i, c in enumerate(regions_class_order):
        seg[probs[i] > 0.5] = float(c)
    return seg


def dice_per_label(seg: torch.Tensor, gt: torch.Tensor, labels):
    """evaluation/evaluator.py + metrics.py `dice`: 2 TP / (2 TP + FP + FN) per label, nan when both are empty."""
    out = OrderedDict()
    for l in labels:
        p, g = seg == float(l), gt == float(l)
        tp = float((p & g).sum())
        den = float(p.sum()) + float(g.sum())
        out[int(l)] = (2.0 * tp / den) if den > 0 else float("nan")
    return out


class ValidationMixin:
    def validate(self, do_mirroring: bool = True, use_sliding_window: bool = True, step_size: float = 0.5,
                 save_softmax: bool = True, use_gaussian: bool = True, overwrite: bool = True,
                 validation_folder_name: str = 'validation_raw', debug: bool = False, all_in_gpu: bool = False,
                 segmentation_export_kwargs: dict = None, run_postprocessing_on_folds: bool = False, export_fn=None):
        """Signature of MT:129-132.  `self.dataset_val`: {case key: {'data': array [c + 1, X, Y, Z] (last channel = label
        map) or 'data_file': npz path, 'properties': dict or 'properties_file': pickle path}}.  Returns
        {dataset: {'cases': {key: {label: dice}}, 'mean': {label: mean dice}}} on every rank."""
        assert self.was_initialized, "must initialize, ideally with checkpoint (or train first)"
        assert getattr(self, "dataset_val", None), "validate needs self.dataset_val (the reference loads it in do_split)"
        net = self.network
        ds, mode = net.do_ds, net.training
        net.do_ds = False
        net.eval()
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank() if world > 1 else 0
        mirror_axes = (0, 1, 2) if do_mirroring else ()       # default_data_augmentation.py:70 (`mirror_axes`)
        tb = list(getattr(self, "transpose_backward", None) or self.plans.get('transpose_backward', [0, 1, 2]))
        all_keys = list(self.dataset_val.keys())
        my_keys = all_keys[rank::world]                       # MT:199-200
        mine = {}
        try:
            for k in my_keys:
                entry = self.dataset_val[k]
                data = entry['data'] if 'data' in entry else np.load(entry['data_file'])['data']
                props = entry.get('properties')
                if props is None and 'properties_file' in entry:
                    with open(entry['properties_file'], 'rb') as f:
                        props = pickle.load(f)
                data = np.array(data, dtype=np.float32, copy=True)
                data[-1][data[-1] == -1] = 0                  # MT:229
                dataset = dataset_name_of_case(k)
                seg_all, probs = net.predict_3D(data[:-1], do_mirroring=do_mirroring, mirror_axes=mirror_axes,
                                                use_sliding_window=use_sliding_window, step_size=step_size,
                                                patch_size=tuple(self.patch_size),
                                                regions_class_order=self.regions_class_order, use_gaussian=use_gaussian,
                                                all_in_gpu=all_in_gpu, verbose=False, return_device_tensors=True)
                probs = probs.permute([0] + [i + 1 for i in tb])                              # MT:237
                chans = [MultiTalent_region_output_idx_mapping[i] for i in MultiTalent_valid_regions[dataset]]  # MT:279
                sel = probs[chans]
                order = MultiTalent_regions_class_order[dataset]                              # MT:282
                seg = assemble_label_map(sel, order)
                gt = torch.from_numpy(data[-1]).to(seg.device).permute(tb)
                labels = (props or {}).get('valid_labels', order)
                mine[k] = (dataset, dice_per_label(seg, gt, labels))
                if export_fn is not None:
                    export_fn(k, sel, seg, props, dict(save_softmax=save_softmax, overwrite=overwrite,
                                                       validation_folder_name=validation_folder_name,
                                                       segmentation_export_kwargs=segmentation_export_kwargs))
        finally:
            net.train(mode)
            net.do_ds = ds
        gathered = [mine]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)            # the reference meets at a barrier, rank 0 scores (MT:300-315)
        summary = {}
        for part in gathered:
            for k, (dataset, scores) in part.items():
                summary.setdefault(dataset, {'cases': OrderedDict()})['cases'][k] = scores
        for dataset, d in summary.items():
            labels = sorted({l for s in d['cases'].values() for l in s})
            d['mean'] = {l: float(np.nanmean([s.get(l, np.nan) for s in d['cases'].values()])) for l in labels}
        self.validation_summary = summary
        return summary
