"""`validate` (MultiTalent_Trainer_DDP.py:129-322, compute half): per-dataset channel selection, in-order label-map
assembly and Dice per label on the device against a numpy restatement of the same steps on the predictor's output."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_validate_scores_equal_numpy_restatement():
    from multitalent_b200.dataset_conversion.Task100_MultiTalent import (MultiTalent_region_output_idx_mapping,
                                                                         MultiTalent_regions_class_order,
                                                                         MultiTalent_valid_regions)
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_case
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    patch = (16, 32, 32)
    plans = default_plans(patch_size=patch, batch_size=2)
    plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
    tr = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False)
    torch.manual_seed(0)
    tr.initialize(True)
    with torch.no_grad():  # decisive heads, so that the thresholded masks are not empty
        for h in tr.network.seg_outputs:
            h.weight.mul_(30.0)
    rng = np.random.RandomState(0)
    cases = {}
    for key, task in (("017_case_a", "Task017_AbdominalOrganSegmentation"), ("003_case_b", "Task003_Liver"),
                      ("017_case_c", "Task017_AbdominalOrganSegmentation")):
        vol, lab = synthetic_case((24, 40, 48), task, rng)
        lab[0, 0, :4] = -1                                      # nnU-Net marks "outside the nonzero mask" with -1
        cases[key] = {'data': np.stack([vol, lab]), 'properties': {}}
    tr.dataset_val = cases
    exported = []
    summary = tr.validate(do_mirroring=False, export_fn=lambda k, probs, seg, props, kw: exported.append((k, tuple(seg.shape))))
    assert [e[0] for e in exported] == list(cases) and exported[0][1] == (24, 40, 48)
    assert set(summary) == {"Task017_AbdominalOrganSegmentation", "Task003_Liver"}
    assert list(summary["Task017_AbdominalOrganSegmentation"]['cases']) == ["017_case_a", "017_case_c"]
    for key, entry in cases.items():
        task = [t for t in MultiTalent_valid_regions if t.startswith("Task" + key[:3])][0]
        _, prob = tr.predict_preprocessed_data_return_seg_and_softmax(entry['data'][:-1], do_mirroring=False, verbose=False)
        chans = [MultiTalent_region_output_idx_mapping[r] for r in MultiTalent_valid_regions[task]]
        order = MultiTalent_regions_class_order[task]
        seg = np.zeros(prob.shape[1:], dtype=np.float32)
        for i, c in enumerate(order):
            seg[prob[chans][i] > 0.5] = c
        gt = entry['data'][-1].copy()
        gt[gt == -1] = 0
        got = summary[task]['cases'][key]
        assert list(got) == list(order)
        some = False
        for l in order:
            p, g = seg == l, gt == l
            den = p.sum() + g.sum()
            want = 2.0 * (p & g).sum() / den if den > 0 else float("nan")
            assert (np.isnan(want) and np.isnan(got[l])) or abs(got[l] - want) < 1e-9, (key, l, got[l], want)
            some = some or (den > 0)
        assert some
    m = summary["Task017_AbdominalOrganSegmentation"]['mean']
    a, c = (summary["Task017_AbdominalOrganSegmentation"]['cases'][k] for k in ("017_case_a", "017_case_c"))
    for l in m:
        assert np.isnan(m[l]) or abs(m[l] - np.nanmean([a[l], c[l]])) < 1e-12
    assert tr.network.do_ds is True and tr.network.training
