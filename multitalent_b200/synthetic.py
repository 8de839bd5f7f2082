"""Synthetic CT patches + 13-dataset label maps for benchmarks and smoke tests.

HU-like volume: blocky background N(-50, 200^2) + ellipsoids with HU in [-100, 300], then the plan's CT normalisation
(clip to [p0.5, p99.5], z-score with the plan's global mean/sd -- nnunet/preprocessing/preprocessing.py:275-283).  Sample b
of rank k belongs to dataset MultiTalent_task_ids[(batch_size*k + b) % 13]; its label map holds that dataset's global
label ids painted as the ellipsoids; valid_regions[b] = MultiTalent_valid_regions[dataset].
"""
import numpy as np

from .dataset_conversion.Task100_MultiTalent import (MultiTalent_task_ids, MultiTalent_task_label_maps,
                                                     MultiTalent_valid_regions)
from .plans import default_plans


def synthetic_case(shape, task, rng, n_blobs=6, plans=None):
    plans = plans or default_plans()
    D, H, W = shape
    coarse = rng.normal(-50.0, 200.0, size=((D + 3) // 4, (H + 3) // 4, (W + 3) // 4)).astype(np.float32)
    vol = np.repeat(np.repeat(np.repeat(coarse, 4, 0), 4, 1), 4, 2)[:D, :H, :W].copy()
    lab = np.zeros(shape, dtype=np.float32)
    labels = MultiTalent_task_label_maps[task][1]
    az, ay, ax = np.arange(D, dtype=np.float32)[:, None, None], np.arange(H, dtype=np.float32)[None, :, None], \
        np.arange(W, dtype=np.float32)[None, None, :]
    for k in range(n_blobs):
        c = [rng.uniform(0.15, 0.85) * s for s in shape]
        r = [max(1.5, rng.uniform(0.08, 0.25) * s) for s in shape]
        m = ((az - c[0]) / r[0]) ** 2 + ((ay - c[1]) / r[1]) ** 2 + ((ax - c[2]) / r[2]) ** 2 <= 1.0
        vol[m] = rng.uniform(-100.0, 300.0)
        lab[m] = labels[k % len(labels)]
    lo, hi = plans['ct_clip']
    vol = (np.clip(vol, lo, hi) - plans['ct_mean']) / plans['ct_sd']
    return vol.astype(np.float32), lab


def synthetic_batch(patch_size, batch_size, rank=0, ds_scales=None, seed=1234, plans=None):
    """-> dict(data [B,1,D,H,W] f32, target list of [B,1,D/s..] f32, properties [{'valid_regions': ...}]) -- the batch
    dictionary the reference's augmenter yields (MultiTalent_Trainer_DDP.py:325-329)."""
    rng = np.random.RandomState(seed + rank)
    vols, labs, props = [], [], []
    for b in range(batch_size):
        task = MultiTalent_task_ids[(batch_size * rank + b) % len(MultiTalent_task_ids)]
        v, l = synthetic_case(tuple(patch_size), task, rng, plans=plans)
        vols.append(v)
        labs.append(l)
        props.append({'valid_regions': MultiTalent_valid_regions[task], 'task': task})
    data = np.stack(vols)[:, None]
    lab = np.stack(labs)[:, None]
    ds_scales = ds_scales or [[1, 1, 1]]
    targets = []
    for s in ds_scales:
        st = [int(round(1 / float(f))) for f in s]
        targets.append(np.ascontiguousarray(lab[..., ::st[0], ::st[1], ::st[2]]))
    return {'data': data, 'target': targets, 'properties': props}
