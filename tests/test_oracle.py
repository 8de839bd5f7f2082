"""CPU tests: pin the oracle against (a) the reference's own known-answer tests for the sliding-window step list,
(b) the committed fixtures generated from the unmodified reference, (c) the live reference when /root/reference exists."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import unet_oracle as O
from oracle import ref_import


# ---- (a) reference KATs: tests/test_steps_for_sliding_window_prediction.py:60-163 -------------------------------
STEP_KATS = [
    ((24, 845, 321), (24, 845, 321), 1, [[0], [0], [0]]),
    ((24, 845, 321), (24, 845, 321), 0.125, [[0], [0], [0]]),
    ((24, 845, 321), (24, 845, 321), 0.5, [[0], [0], [0]]),
    ((123, 143), (123, 143), 1, [[0], [0]]),
    ((123, 143), (123, 143), 0.125, [[0], [0]]),
    ((123, 143), (123, 143), 0.5, [[0], [0]]),
    ((128, 260), (64, 130), 0.5, [[0, 32, 64], [0, 65, 130]]),
    ((128, 260), (64, 130), 0.85, [[0, 32, 64], [0, 65, 130]]),
    ((128, 260), (64, 130), 1, [[0, 64], [0, 130]]),
    ((146, 176, 148), (128, 128, 128), 0.5, [[0, 18], [0, 48], [0, 20]]),
    ((130, 320, 244), (80, 192, 160), 0.5, [[0, 25, 50], [0, 64, 128], [0, 42, 84]]),
    ((130, 320, 244), (80, 192, 160), 0.75, [[0, 50], [0, 128], [0, 84]]),
    ((424, 456, 456), (128, 128, 128), 0.5, [[0, 59, 118, 178, 237, 296], [0, 55, 109, 164, 219, 273, 328],
                                             [0, 55, 109, 164, 219, 273, 328]]),
    ((40, 56, 40), (40, 56, 40), 0.5, [[0], [0], [0]]),
    ((94, 308, 308), (64, 192, 192), 0.5, [[0, 30], [0, 58, 116], [0, 58, 116]]),
]


def _step_fns():
    from multitalent_b200.network_architecture.neural_network import SegmentationNetwork
    return [O.compute_steps_for_sliding_window, SegmentationNetwork._compute_steps_for_sliding_window]


@pytest.mark.parametrize("image,patch,step,expected", STEP_KATS)
def test_steps_known_answers(image, patch, step, expected):
    for fn in _step_fns():
        assert fn(patch, image, step) == expected


def test_steps_properties():
    """The 5000-case property test of the reference (:25-58, :165-181)."""
    rng = np.random.RandomState(0)
    for _ in range(5000):
        dim = rng.choice((2, 3))
        patch = tuple(int(v) for v in rng.randint(16, 1024, dim))
        image = tuple(max(int(rng.randint(p // 2, p * 10)), p) for p in patch)
        step = float(rng.uniform(0.01, 1))
        for fn in _step_fns():
            steps = fn(patch, image, step)
            target = [i * step for i in patch]
            nsteps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image, target, patch)]
            for d in range(dim):
                s = steps[d]
                assert len(s) == nsteps[d] and s[0] == 0 and s[-1] + patch[d] == image[d]
                assert all(s[i + 1] <= s[i] + patch[d] for i in range(len(s) - 1))
                assert all(s[i] + np.ceil(target[d]) >= s[i + 1] for i in range(len(s) - 1))


def test_steps_benchmark_volume():
    """512^3, patch 192x160x128, step .5 -> 5 x 6 x 7 = 210 tiles."""
    s = O.compute_steps_for_sliding_window((192, 160, 128), (512, 512, 512), 0.5)
    assert s[0] == [0, 80, 160, 240, 320] and s[1] == [0, 70, 141, 211, 282, 352] and len(s[2]) == 7


# ---- (b) fixtures ---------------------------------------------------------------------------------------------------
def test_tables_match_fixture():
    with open(os.path.join(GOLD, "tables.json")) as f:
        t = json.load(f)
    from multitalent_b200.dataset_conversion import Task100_MultiTalent as P
    for mod_ids, regions, chan, valid, lmaps in (
            (O.TASK_IDS, O.REGIONS, O.REGION_CHANNEL, O.VALID_REGIONS, O.TASK_LABEL_MAPS),
            (P.MultiTalent_task_ids, P.MultiTalent_regions, P.MultiTalent_region_output_idx_mapping,
             P.MultiTalent_valid_regions, P.MultiTalent_task_label_maps)):
        assert list(mod_ids) == t["task_ids"]
        assert [[k, list(v)] for k, v in regions.items()] == t["regions"]
        assert dict(chan) == t["region_output_idx_mapping"]
        assert {k: list(v) for k, v in valid.items()} == t["valid_regions"]
        assert {k: [list(v[0]), list(v[1])] for k, v in lmaps.items()} == t["task_label_maps"]
    assert {str(k): v for k, v in P.MultiTalent_labels.items()} == t["labels"]
    assert {k: list(v) for k, v in P.MultiTalent_regions_class_order.items()} == t["regions_class_order"]


def _sd(blob):
    return {k[len("param/"):]: torch.from_numpy(v) for k, v in blob.items() if k.startswith("param/")}


def test_oracle_forward_loss_grads_match_fixture(golden_small):
    blob, meta = golden_small
    torch.set_num_threads(1)
    sd = {k: v.clone().requires_grad_(True) for k, v in _sd(blob).items()}
    x = torch.from_numpy(blob["x"])
    out = O.generic_unet_forward(x, sd, meta["pool"], meta["convk"])
    for i, o in enumerate(out):
        np.testing.assert_allclose(o.detach().numpy(), blob["logits_%d" % i], rtol=0, atol=2e-5)
    tg = [torch.from_numpy(blob["target_%d" % i]) for i in range(len(out))]
    l, ce, dc = O.multitalent_loss(out, tg, meta["valid_regions"], blob["ds_loss_weights"])
    np.testing.assert_allclose([l.item(), ce.item(), dc.item()], blob["loss"], rtol=1e-5)
    l.backward()
    for k, v in sd.items():
        g = blob["grad/" + k]
        np.testing.assert_allclose(v.grad.numpy(), g, rtol=0, atol=1e-5 + 1e-3 * np.abs(g).max())


def test_oracle_ddp_semantics_single_rank_identity(golden_small):
    """multitalent_loss_ddp with world 1 and no other ranks == multitalent_loss."""
    blob, meta = golden_small
    out = [torch.from_numpy(blob["logits_%d" % i]).requires_grad_(True) for i in range(3)]
    tg = [torch.from_numpy(blob["target_%d" % i]) for i in range(3)]
    a = O.multitalent_loss(out, tg, meta["valid_regions"], blob["ds_loss_weights"])
    b = O.multitalent_loss_ddp(out, tg, meta["valid_regions"], blob["ds_loss_weights"], None, 1)
    assert abs(a[0].item() - b[0].item()) < 1e-6


def test_oracle_sliding_window_matches_fixture(golden_small, golden_sliding):
    blob, meta = golden_small
    sd = _sd(blob)
    torch.set_num_threads(2)

    def net_fn(t):
        with torch.no_grad():
            return torch.sigmoid(O.generic_unet_forward(t, sd, meta["pool"], meta["convk"], do_ds=False))

    vol = golden_sliding["vol"]
    patch = (8, 16, 16)
    seg, prob = O.predict_3d_tiled(net_fn, vol[None], patch, 47, 0.5, True, (0, 1, 2), True, tuple(range(47)))
    np.testing.assert_allclose(prob[:, ::2, ::2, ::2], golden_sliding["prob_mirror_sub"], atol=2e-5)
    assert (seg != golden_sliding["seg_mirror"]).mean() < 1e-3
    seg, prob = O.predict_3d_tiled(net_fn, vol[None], patch, 47, 0.5, False, (0, 1, 2), True, tuple(range(47)))
    np.testing.assert_allclose(prob[:, ::2, ::2, ::2], golden_sliding["prob_nomirror_sub"], atol=2e-5)


def test_gaussian_and_pad():
    g = O.get_gaussian((8, 16, 16))
    assert g.dtype == np.float32 and g.max() == 1.0 and g.min() > 0
    from multitalent_b200.network_architecture.neural_network import SegmentationNetwork, pad_nd_image
    np.testing.assert_array_equal(g, SegmentationNetwork._get_gaussian((8, 16, 16)))
    x = np.arange(2 * 5 * 7 * 9, dtype=np.float32).reshape(2, 5, 7, 9)
    for fn in (O.pad_nd_image, pad_nd_image):
        y, sl = fn(x, (8, 6, 12), "constant", {'constant_values': 0}, True, None)
        assert y.shape == (2, 8, 7, 12)
        np.testing.assert_array_equal(y[tuple(sl)], x)
        assert sl[1] == slice(1, 6) and sl[3] == slice(1, 10)
        y2, _ = fn(x, None, "constant", None, True, [4, 4, 4])
        assert y2.shape == (2, 8, 8, 12)


def test_clip_and_sgd_matches_torch():
    torch.manual_seed(0)
    ps = [torch.randn(5, 3), torch.randn(7)]
    gs = [torch.randn(5, 3) * 10, torch.randn(7) * 10]
    ref_p = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.SGD(ref_p, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    bufs = [None, None]
    cur = [p.clone() for p in ps]
    for step in range(3):
        for p, g in zip(ref_p, gs):
            p.grad = g.clone() * (step + 1)
        torch.nn.utils.clip_grad_norm_(ref_p, 12)
        opt.step()
        cur, bufs, _ = O.clip_and_sgd_step(cur, [g * (step + 1) for g in gs], bufs, 1e-2)
        for a, b in zip(cur, ref_p):
            np.testing.assert_allclose(a.numpy(), b.detach().numpy(), atol=1e-6)


# ---- (c) live reference ----------------------------------------------------------------------------------------------
@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference():
    ref_import.install()
    pool, ck = [[2, 2, 2], [2, 2, 2], [1, 2, 2]], [[3, 3, 3]] * 4
    net = ref_import.build_reference_generic_unet(1, 8, 47, pool, ck, seed=3)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    x = torch.randn(2, 1, 8, 16, 16)
    with torch.no_grad():
        r = net(x)
        o = O.generic_unet_forward(x, sd, pool, ck)
    for a, b in zip(r, o):
        assert float((a - b).abs().max()) < 1e-6
    rng = np.random.RandomState(5)
    tasks = ["Task046_AbdOrgSegm2", "Task064_KiTS_labelsFixed"]
    lab = np.stack([O.synthetic_ct_and_labels((8, 16, 16), t, rng)[1] for t in tasks])[:, None]
    tg = [torch.from_numpy(a) for a in O.downsample_targets(lab, [[1, 1, 1], [.5, .5, .5], [.25, .25, .25]])]
    valid = [O.VALID_REGIONS[t] for t in tasks]
    w = O.multitalent_ds_loss_weights(3)
    la = ref_import.reference_compute_loss([a.clone() for a in r], tg, valid, w)
    lb = O.multitalent_loss(list(r), tg, valid, w)
    for a, b in zip(la, lb):
        assert abs(a.item() - b.item()) < 1e-5
