"""patch sampling + crop/pad, deep-supervision target down-sampling, export resampling.
not-gpu: the host logic (same np.random consumption, bbox arithmetic, oversampling) against a fixture drawn from the
UNMODIFIED reference `DataLoader3D` (oracle/make_golden_dataloader.py) with the oracle's numpy crop swapped in, and the
oracle's resize restatements against their defining formulas.  gpu: the device kernels against the same fixture / scipy."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import unet_oracle as O

CFG = dict(patch_size=(14, 20, 20), final_patch_size=(10, 16, 16), batch_size=3, oversample=0.34, pad_sides=(2, 0, 4), seed=7)


@pytest.fixture(scope="module")
def dl_gold():
    return dict(np.load(os.path.join(GOLD, "dataloader_small.npz"), allow_pickle=False))


def _dataset(blob, device):
    keys = [k[len("case/"):] for k in blob if k.startswith("case/")]
    ds = {}
    for k in keys:
        cl = {l: blob["loc/%s/%d" % (k, l)] for l in (1, 2)}
        arr = blob["case/" + k]
        ds[k] = {'data': torch.from_numpy(arr).to(device) if device else arr, 'properties': {'class_locations': cl}}
    return ds


def _loader(blob, device):
    from multitalent_b200.training.dataloading.dataset_loading import DataLoader3D
    return DataLoader3D(_dataset(blob, device), CFG["patch_size"], CFG["final_patch_size"], CFG["batch_size"], False,
                        oversample_foreground_percent=CFG["oversample"], pad_mode="constant", pad_sides=CFG["pad_sides"],
                        memmap_mode='r', sampling_probabilities=blob["probs"])


def test_patch_sampling_host_logic_matches_reference_fixture(dl_gold):
    """Same keys, same boxes, same padding as the reference for the same np.random seed (numpy crop from the oracle)."""
    dl = _loader(dl_gold, "cpu")

    def np_crop(case, bbox, patch, pad_mode, pad_kwargs, out_data, out_seg):
        d, s = O.crop_and_pad_case(case.numpy(), bbox, patch, pad_mode, pad_kwargs)
        out_data.copy_(torch.from_numpy(d))
        out_seg.copy_(torch.from_numpy(s))
    dl._crop = np_crop
    np.random.seed(CFG["seed"])
    for i in range(3):
        b = dl.generate_train_batch()
        assert [str(k) for k in b['keys']] == [str(k) for k in dl_gold["batch%d/keys" % i]]
        assert np.array_equal(b['data'].numpy(), dl_gold["batch%d/data" % i])
        assert np.array_equal(b['seg'].numpy(), dl_gold["batch%d/seg" % i])
    assert dl.get_do_oversample(2) and not dl.get_do_oversample(1)     # round(3 * 0.66) = 2 -> only the last sample


def test_nearest_resize_formula_equals_scipy_zoom():
    """The kernel's integer formula floor((2 o + 1) in / (2 out)) == what scipy's grid-mode zoom (= skimage resize,
    order 0) samples, for the deep-supervision ratios and for ragged ones."""
    rng = np.random.RandomState(0)
    for n_in, n_out in [(16, 8), (32, 8), (48, 3), (17, 9), (20, 7), (12, 12), (9, 4), (64, 4)]:
        v = rng.permutation(n_in).astype(np.float64)
        ref = O.resize_like_skimage(v, (n_out,), 0)
        idx = np.clip(((2 * np.arange(n_out) + 1) * n_in) // (2 * n_out), 0, n_in - 1)
        assert np.array_equal(ref, v[idx]), (n_in, n_out)


def test_oracle_ds_targets_shapes():
    seg = np.random.RandomState(1).randint(0, 5, size=(2, 1, 16, 32, 32)).astype(np.float32)
    out = O.downsample_seg_for_ds(seg, [[1, 1, 1], [0.5, 0.5, 0.5], [0.25, 0.25, 0.25], [0.125, 0.25, 0.25]])
    assert out[0] is seg and [o.shape[2:] for o in out[1:]] == [(8, 16, 16), (4, 8, 8), (2, 8, 8)]
    assert np.array_equal(out[1], seg[:, :, 1::2, 1::2, 1::2])          # factor 2 samples the odd voxels


@pytest.mark.gpu
def test_device_patch_loader_matches_reference_fixture(dl_gold):
    dl = _loader(dl_gold, "cuda")
    np.random.seed(CFG["seed"])
    for i in range(3):
        b = dl.generate_train_batch()
        assert b['data'].is_cuda and [str(k) for k in b['keys']] == [str(k) for k in dl_gold["batch%d/keys" % i]]
        assert np.array_equal(b['data'].cpu().numpy(), dl_gold["batch%d/data" % i])
        assert np.array_equal(b['seg'].cpu().numpy(), dl_gold["batch%d/seg" % i])


@pytest.mark.gpu
def test_device_crop_pad_edge_mode_equals_numpy():
    from multitalent_b200.training.dataloading.dataset_loading import crop_and_pad_case
    rng = np.random.RandomState(2)
    case = rng.randn(3, 11, 9, 13).astype(np.float32)
    for lb in [(-3, -2, -5), (4, 2, 6), (0, 0, 0), (-20, 3, 1)]:
        d, s = crop_and_pad_case(torch.from_numpy(case).cuda(), lb, (12, 10, 12), pad_mode="edge")
        rd, rs = O.crop_and_pad_case(case, lb, (12, 10, 12), "edge") if all(l > -11 for l in lb) else (None, None)
        if rd is not None:
            assert np.array_equal(d.cpu().numpy(), rd) and np.array_equal(s.cpu().numpy(), rs)
        else:  # box entirely outside along x: np.pad('edge') cannot pad an empty array, the kernel clamps
            assert np.array_equal(d.cpu().numpy()[:, 0], d.cpu().numpy()[:, -1])


@pytest.mark.gpu
def test_device_ds_targets_equal_oracle():
    from multitalent_b200.training.data_augmentation.downsampling import DownsampleSegForDSTransform2
    seg = np.random.RandomState(4).randint(-1, 48, size=(2, 1, 32, 40, 48)).astype(np.float32)
    scales = [[1, 1, 1], [0.5, 0.5, 0.5], [0.25, 0.25, 0.25], [0.125, 0.125, 0.125], [1 / 16, 0.125, 0.0625]]
    tr = DownsampleSegForDSTransform2(scales, 0, input_key="seg", output_key="target")
    got = tr(seg=torch.from_numpy(seg).cuda())["target"]
    ref = O.downsample_seg_for_ds(seg, scales)
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert np.array_equal(g.cpu().numpy(), r)


@pytest.mark.gpu
@pytest.mark.parametrize("spacing,want_sep", [((1.0, 1.0, 1.0), False), ((5.0, 0.8, 0.8), True)])
def test_device_export_resampling_equals_oracle(spacing, want_sep):
    from multitalent_b200.inference.segmentation_export import resample_softmax_and_threshold
    rng = np.random.RandomState(5)
    C, cur, after = 5, (12, 20, 18), (17, 31, 25)
    prob = rng.rand(C, *cur).astype(np.float32)
    order_cls = (3, 1, 4, 2, 9)
    props = {'size_after_cropping': after, 'original_size_of_raw_data': (20, 33, 30), 'original_spacing': spacing,
             'spacing_after_resampling': (1.5, 1.0, 1.0), 'crop_bbox': [[2, 19], [1, 32], [3, 28]]}
    seg, pr = resample_softmax_and_threshold(torch.from_numpy(prob).cuda(), props, 1, order_cls, None, 0,
                                             return_probabilities=True)
    orders = [0, 1, 1] if want_sep else [1, 1, 1]
    rseg, rprob = O.resample_and_threshold(prob, after, orders, order_cls)
    assert pr.dtype == torch.float16 and tuple(pr.shape) == (C,) + after
    assert float(np.abs(pr.float().cpu().numpy() - rprob).max()) < 2e-3          # fp16 storage of values in [0, 1]
    full = np.zeros((20, 33, 30), dtype=np.uint8)
    full[2:19, 1:32, 3:28] = rseg
    got = seg.cpu().numpy()
    assert got.shape == full.shape
    mism = got != full
    # a voxel may only differ where some probability sits within fp32 rounding of the threshold
    assert mism.mean() < 1e-3
    if mism.any():
        near = (np.abs(rprob - 0.5) < 1e-5).any(0)
        fn = np.zeros_like(mism)
        fn[2:19, 1:32, 3:28] = near
        assert not (mism & ~fn).any()
    # argmax mode (softmax networks, region_class_order None)
    seg2, _ = resample_softmax_and_threshold(torch.from_numpy(prob).cuda(), dict(props, crop_bbox=None), 1, None, False)
    r2, _ = O.resample_and_threshold(prob, after, [1, 1, 1], None)
    assert (seg2.cpu().numpy() != r2).mean() < 1e-3
