"""Online hard-Dice evaluation against the reference's own method
(MultiTalent_Trainer_DDP.py:372-430), on the CPU: the implementation is device-agnostic torch code."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multitalent_b200.training.online_evaluation import OnlineEvaluationMixin, hard_tp_fp_fn  # noqa: E402
from oracle import ref_import, unet_oracle as O  # noqa: E402


def _case(seed, shape=(6, 10, 12), tasks=("Task046_AbdOrgSegm2", "Task064_KiTS_labelsFixed", "Task003_Liver")):
    rng = np.random.RandomState(seed)
    lab = np.stack([O.synthetic_ct_and_labels(shape, t, rng)[1] for t in tasks])[:, None].astype(np.float32)
    logits = torch.from_numpy(rng.randn(len(tasks), 47, *shape).astype(np.float32))
    # make the predictions correlate with the labels so that tp is not trivially small
    for b, t in enumerate(tasks):
        for r in O.VALID_REGIONS[t]:
            j = O.REGION_INDEX[r] if hasattr(O, "REGION_INDEX") else list(O.REGIONS.keys()).index(r)
            m = np.isin(lab[b, 0], list(O.REGIONS[r]))
            logits[b, j][torch.from_numpy(m)] += 1.5
    return logits, torch.from_numpy(lab), [O.VALID_REGIONS[t] for t in tasks]


def _brute_force(logits, target, valid):
    B, C = logits.shape[:2]
    tp, fp, fn = np.zeros((B, C)), np.zeros((B, C)), np.zeros((B, C))
    names = list(O.REGIONS.keys())
    for b in range(B):
        for r in valid[b]:
            j = names.index(r)
            gt = np.isin(target[b, 0].numpy(), list(O.REGIONS[r]))
            pr = (torch.sigmoid(logits[b, j]) > 0.5).numpy()
            tp[b, j], fp[b, j], fn[b, j] = (pr & gt).sum(), (pr & ~gt).sum(), (~pr & gt).sum()
    return tp, fp, fn


def test_counts_match_the_definition():
    logits, target, valid = _case(0)
    tp, fp, fn = hard_tp_fp_fn(logits, target, valid)
    btp, bfp, bfn = _brute_force(logits, target, valid)
    np.testing.assert_array_equal(tp.numpy(), btp)
    np.testing.assert_array_equal(fp.numpy(), bfp)
    np.testing.assert_array_equal(fn.numpy(), bfn)
    assert tp.sum() > 0 and fp.sum() > 0 and fn.sum() > 0
    # channels a sample's dataset does not label stay exactly zero
    names = list(O.REGIONS.keys())
    for b, v in enumerate(valid):
        off = [j for j in range(47) if names[j] not in v]
        assert float(tp[b, off].abs().sum() + fp[b, off].abs().sum() + fn[b, off].abs().sum()) == 0.0


def test_strided_channels_last_logits_give_the_same_counts():
    """The trainer hands over NCDHW-shaped VIEWS of NDHWC buffers."""
    logits, target, valid = _case(1)
    view = logits.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    assert not view.is_contiguous()
    for a, b in zip(hard_tp_fp_fn(view, target, valid), hard_tp_fp_fn(logits, target, valid)):
        assert torch.equal(a, b)


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree only exists in the build container")
def test_matches_the_reference_methods():
    ref_import.install()
    ref_import.init_gloo_single()
    from nnunet.training.network_training.custom_trainers.MultiTalent.MultiTalent.MultiTalent_Trainer_DDP import \
        MultiTalent_trainer_ddp as Ref

    def fresh():
        return SimpleNamespace(online_eval_foreground_dc=[], online_eval_tp=[], online_eval_fp=[], online_eval_fn=[],
                               all_val_eval_metrics=[], print_to_log_file=lambda *a, **k: None)
    ref = fresh()

    class Mine(OnlineEvaluationMixin):
        pass
    mine = Mine()
    for seed in (2, 3, 4):
        logits, target, valid = _case(seed)
        Ref.run_online_evaluation(ref, [logits], [target], valid)
        mine.run_online_evaluation([logits], [target], valid)
    for name in ("online_eval_foreground_dc", "online_eval_tp", "online_eval_fp", "online_eval_fn"):
        np.testing.assert_allclose(np.array(getattr(mine, name), dtype=np.float64),
                                   np.array(getattr(ref, name), dtype=np.float64), rtol=1e-6, atol=1e-7, err_msg=name)
    Ref.finish_online_evaluation(ref)
    per_class = mine.finish_online_evaluation()
    assert mine.all_val_eval_metrics[-1] == pytest.approx(ref.all_val_eval_metrics[-1], rel=1e-6)
    # reference quirk kept: the accumulators are [B, 47] per iteration, so the "per class" list has one row per batch slot
    assert len(per_class) == 3 and all(len(r) == 47 for r in per_class)
    assert mine.online_eval_tp == [] and ref.online_eval_tp == []


@pytest.mark.gpu
def test_fused_hard_counts_from_the_loss_kernel_equal_the_standalone_evaluation(monkeypatch):
    """Inside a trainer step the hard tp / fp / fn come out of `mtb200_mt_loss_stats` (the validation iteration reads the
    logits once for loss and evaluation): identical to the stand-alone boolean reductions, fp32 and bf16 storage, and
    the stand-alone pass is really not run."""
    import multitalent_b200.training.online_evaluation as OE
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_batch
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    patch = (16, 32, 32)
    plans = default_plans(patch_size=patch, batch_size=3)
    plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
    for dtype in (torch.float32, torch.bfloat16):
        tr = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False, native_dtype=dtype)
        torch.manual_seed(0)
        tr.initialize(True)
        batch = synthetic_batch(patch, 3, 2, tr.deep_supervision_scales)
        valid = [p['valid_regions'] for p in batch['properties']]
        # stand-alone counts on the same logits
        with torch.no_grad():
            out = tr.network(torch.from_numpy(batch['data']).cuda())
        tp, fp, fn = OE.hard_tp_fp_fn(out[0], torch.from_numpy(batch['target'][0]).cuda(), valid)
        assert float(tp.sum()) + float(fp.sum()) > 0 and float(fn.sum()) + float(tp.sum()) > 0

        def boom(*a, **k):
            raise AssertionError("the stand-alone evaluation pass must not run inside a trainer step")
        monkeypatch.setattr(OE, "hard_tp_fp_fn", boom)
        tr.run_iteration(iter([batch]), do_backprop=False, run_online_evaluation=True)
        monkeypatch.undo()
        # the reference's accumulators hold [B, 47] per iteration (sum over the RANK axis only, MT:404-407)
        assert np.array_equal(np.array(tr.online_eval_tp[0]), tp.cpu().numpy())
        assert np.array_equal(np.array(tr.online_eval_fp[0]), fp.cpu().numpy())
        assert np.array_equal(np.array(tr.online_eval_fn[0]), fn.cpu().numpy())
        per_class = tr.finish_online_evaluation()
        # the reference's quirk: the accumulators are [B, 47], so the 'per class' list has one row per batch index
        assert len(per_class) == 3 and len(per_class[0]) == 47 and len(tr.all_val_eval_metrics) >= 1
