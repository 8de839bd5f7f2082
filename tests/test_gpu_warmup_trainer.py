"""nnUNetTrainerV2_warmupsegheads on the native kernels: the heads-only phase against the CPU
oracle (loss value, head gradients through one AdamW step), frozen-trunk path == full-backward path for the heads, the
switch to whole-network SGD, and the pretrained-weight transfer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PATCH = (16, 32, 32)


def _plans():
    from multitalent_b200.plans import default_plans
    plans = default_plans(patch_size=PATCH, batch_size=2)
    plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
    plans['base_num_features'] = 8
    plans['num_classes'] = 3  # + background = 4 output channels
    return plans


def _batch(scales, seed=0):
    rng = np.random.RandomState(seed)
    data = rng.randn(2, 1, *PATCH).astype(np.float32)
    lab = rng.randint(0, 4, size=(2, 1) + PATCH).astype(np.float32)
    tg = []
    for s in scales:
        st = [int(round(1 / v)) for v in s]
        tg.append(np.ascontiguousarray(lab[:, :, ::st[0], ::st[1], ::st[2]]))
    return {'data': data, 'target': tg}


def _trainer(freeze):
    from multitalent_b200.training.network_training.nnUNetTrainerV2_warmup import nnUNetTrainerV2_warmupsegheads
    tr = nnUNetTrainerV2_warmupsegheads(_plans(), 0, freeze_trunk_during_head_warmup=freeze)
    torch.manual_seed(0)
    tr.initialize(True)
    return tr


def test_heads_only_phase_matches_oracle_and_frozen_trunk_path():
    from oracle import unet_oracle as O
    frozen, full = _trainer(True), _trainer(False)
    assert frozen.num_classes == 4 and frozen.seg_heads_only and frozen.lr == pytest.approx(5e-5)
    assert not any(p.requires_grad for n, p in frozen.network.named_parameters() if not n.startswith("seg_outputs."))
    assert all(p.requires_grad for p in full.network.parameters())
    batch = _batch(frozen.deep_supervision_scales)
    sd0 = {k: v.detach().cpu().clone() for k, v in frozen.network.state_dict().items()}
    assert all(torch.equal(v.cpu(), sd0[k]) for k, v in full.network.state_dict().items())

    # oracle: forward + softmax DC+CE with deep supervision + gradients of the heads, then torch's AdamW on the heads
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    out = O.generic_unet_forward(torch.from_numpy(batch['data']), sdg, frozen.net_num_pool_op_kernel_sizes,
                                 frozen.net_conv_kernel_sizes)
    lo = O.dc_ce_loss(out, [torch.from_numpy(t) for t in batch['target']], frozen.ds_loss_weights)
    lo.backward()
    heads = [k for k in sd0 if k.startswith("seg_outputs.")]
    # the lowest-resolution output has deep-supervision weight 0: no gradient in the reference, exact zeros here
    og = {k: (sdg[k].grad if sdg[k].grad is not None else torch.zeros_like(sd0[k])) for k in heads}
    assert float(torch.sqrt(sum((g.double() ** 2).sum() for g in og.values()))) < 12.0  # the clip is a no-op here

    l_frozen = float(frozen.run_iteration(iter([batch]), True))
    l_full = float(full.run_iteration(iter([batch]), True))
    assert l_frozen == pytest.approx(float(lo), rel=1e-4, abs=1e-4) and l_full == pytest.approx(l_frozen, rel=1e-6)
    gf = {n: p.grad.detach().cpu() for n, p in frozen.network.named_parameters() if n.startswith("seg_outputs.")}
    gu = {n: p.grad.detach().cpu() for n, p in full.network.named_parameters() if n.startswith("seg_outputs.")}
    gmax = max(float(g.abs().max()) for g in og.values())
    for k in heads:  # head gradients: frozen-trunk path == full backward == oracle
        assert float((gf[k] - og[k]).abs().max()) <= 5e-3 * gmax + 1e-9, k
        assert float((gf[k] - gu[k]).abs().max()) <= 1e-5 * gmax + 1e-12, k
    new_frozen, new_full = frozen.network.state_dict(), full.network.state_dict()
    moved = 0
    for k in sd0:
        if k.startswith("seg_outputs."):
            moved += int(not torch.equal(new_frozen[k].cpu(), sd0[k]))
            # AdamW's first step is ~lr * sign(g) per entry (lr = 5e-5): same gradients => same step
            assert float((new_frozen[k].cpu() - new_full[k].cpu()).abs().max()) <= 5e-6, k
            assert float((new_frozen[k].cpu() - sd0[k]).abs().max()) <= 1.01 * 5e-5 + 1e-7, k
        else:  # the trunk does not move in this phase, frozen or not (the optimizer only holds the heads)
            assert torch.equal(new_frozen[k].cpu(), sd0[k]) and torch.equal(new_full[k].cpu(), sd0[k]), k
    assert moved >= 2
    assert all(p.grad is None for n, p in frozen.network.named_parameters() if not n.startswith("seg_outputs."))
    assert all(p.grad is not None for p in full.network.parameters())


def test_switch_to_whole_network_sgd_and_schedule():
    tr = _trainer(True)
    batch = _batch(tr.deep_supervision_scales, seed=1)
    tr.epoch = tr.warmup_duration            # end of the last heads-only epoch
    assert tr.on_epoch_end()
    assert not tr.seg_heads_only and isinstance(tr.optimizer, torch.optim.SGD)
    assert all(p.requires_grad for p in tr.network.parameters())
    # maybe_update_lr ran with self.epoch == 10 (network_trainer.py:609), then the counter moved on (:490)
    assert tr.epoch == 11 and tr.lr == pytest.approx(1 / 50 * 1e-2)
    before = {k: v.detach().clone() for k, v in tr.network.state_dict().items()}
    l0 = float(tr.run_iteration(iter([batch]), True))
    moved = [k for k, v in tr.network.state_dict().items() if not torch.equal(v, before[k])]
    assert any(k.startswith("conv_blocks_context.0") for k in moved) and any(k.startswith("seg_outputs") for k in moved)
    losses = [float(tr.run_iteration(iter([batch]), True)) for _ in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < l0


def test_pretrained_trunk_transfer_into_the_native_network():
    from multitalent_b200.plans import default_plans
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    plans = _plans()
    mt_plans = default_plans(patch_size=PATCH, batch_size=2)
    for k in ('pool_op_kernel_sizes', 'conv_kernel_sizes'):
        mt_plans['plans_per_stage'][1][k] = plans['plans_per_stage'][1][k]
    mt_plans['base_num_features'] = 8
    mt = MultiTalent_trainer_ddp(mt_plans, 0, 0, init_distributed=False)
    torch.manual_seed(5)
    mt.initialize(True)
    ckpt = {'state_dict': {"module." + k: v.detach().cpu().clone() for k, v in mt.network.state_dict().items()}}
    tr = _trainer(True)
    heads_before = {k: v.detach().clone() for k, v in tr.network.state_dict().items() if k.startswith("seg_outputs.")}
    keys = tr.load_pretrained_weights(ckpt)
    sd = tr.network.state_dict()
    for k, v in mt.network.state_dict().items():
        if k.startswith("seg_outputs."):
            assert k not in keys and torch.equal(sd[k], heads_before[k])   # 47 vs 4 classes
        else:
            assert torch.equal(sd[k].cpu(), v.cpu()), k
    assert np.isfinite(float(tr.run_iteration(iter([_batch(tr.deep_supervision_scales)]), True)))


def test_softmax_sliding_window_matches_oracle():
    """The fine-tuning trainers predict with a channel softmax (`softmax_helper`): the aggregation kernel's softmax mode
    against the oracle's tiled predictor (Gaussian weighting, argmax segmentation), with and without mirroring."""
    from oracle import unet_oracle as O
    tr = _trainer(True)
    sd = {k: v.detach().cpu() for k, v in tr.network.state_dict().items()}
    vol = np.random.RandomState(3).randn(1, 24, 40, 48).astype(np.float32)

    def net_fn(t):
        with torch.no_grad():
            return torch.softmax(O.generic_unet_forward(t, sd, tr.net_num_pool_op_kernel_sizes, tr.net_conv_kernel_sizes,
                                                        do_ds=False), 1)
    for mirror in (False, True):
        seg, prob = tr.predict_preprocessed_data_return_seg_and_softmax(vol, do_mirroring=mirror, verbose=False)
        seg_ref, prob_ref = O.predict_3d_tiled(net_fn, vol, PATCH, 4, 0.5, mirror, (0, 1, 2), True, None)
        assert prob.shape == prob_ref.shape == (4, 24, 40, 48)
        assert float(np.abs(prob - prob_ref).max()) < 1e-3
        np.testing.assert_allclose(prob.sum(0), 1.0, atol=1e-4)
        # argmax may only differ where the two best classes are within the tolerance
        diff = seg != seg_ref
        if diff.any():
            top2 = np.sort(prob_ref, axis=0)[-2:]
            assert float((top2[1] - top2[0])[diff].max()) < 2e-3


def _resenc_plans():
    from multitalent_b200.plans import default_plans
    plans = default_plans("resenc", patch_size=PATCH, batch_size=2)
    sp = plans['plans_per_stage'][1]
    sp['pool_op_kernel_sizes'] = [[1, 1, 1], [1, 2, 2], [2, 2, 2], [2, 2, 2]]
    sp['conv_kernel_sizes'] = [[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]]
    sp['num_blocks_encoder'], sp['num_blocks_decoder'] = (1, 2, 2, 2), (1, 1, 1)
    plans['base_num_features'] = 8
    plans['num_classes'] = 3
    return plans


def test_resenc_warmup_trainer_heads_only_then_whole_network():
    """nnUNetTrainerV2_warmupsegheads_resenc (nnUNetTrainerV2_warmup.py:441-560): FabiansUNet, heads =
    decoder.deep_supervision_outputs, frozen-trunk path == full backward for the heads, loss == oracle, then SGD."""
    from multitalent_b200.training.network_training.nnUNetTrainerV2_warmup import nnUNetTrainerV2_warmupsegheads_resenc
    from oracle import unet_oracle as O
    trs = []
    for freeze in (True, False):
        tr = nnUNetTrainerV2_warmupsegheads_resenc(_resenc_plans(), 0, freeze_trunk_during_head_warmup=freeze)
        torch.manual_seed(0)
        tr.initialize(True)
        with torch.no_grad():   # norm2 is zero-initialised (init_last_bn_before_add_to_0): un-hide conv2 / norm2
            g = torch.Generator().manual_seed(3)
            for n, p in tr.network.named_parameters():
                if n.endswith("norm2.weight"):
                    p.copy_((0.5 + torch.rand(p.shape, generator=g)).to(p.device))
        trs.append(tr)
    frozen, full = trs
    assert frozen.deep_supervision_scales == [[1, 1, 1], [1.0, 0.5, 0.5], [0.5, 0.25, 0.25]]
    assert isinstance(frozen.optimizer, torch.optim.AdamW) and frozen.seg_heads_only
    hp = "decoder.deep_supervision_outputs."
    assert not any(p.requires_grad for n, p in frozen.network.named_parameters() if not n.startswith(hp))
    batch = _batch(frozen.deep_supervision_scales)
    sd0 = {k: v.detach().cpu().clone() for k, v in frozen.network.state_dict().items()}
    sp = frozen.plans['plans_per_stage'][1]
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd0.items() if ".all." not in k}
    out = O.fabians_unet_forward(torch.from_numpy(batch['data']), sdg, sp['num_blocks_encoder'],
                                 sp['pool_op_kernel_sizes'], sp['conv_kernel_sizes'], sp['num_blocks_decoder'])
    lo = O.dc_ce_loss(out, [torch.from_numpy(t) for t in batch['target']], frozen.ds_loss_weights)
    lo.backward()
    l_frozen = float(frozen.run_iteration(iter([batch]), True))
    l_full = float(full.run_iteration(iter([batch]), True))
    assert l_frozen == pytest.approx(float(lo), rel=1e-4, abs=1e-4) and l_full == pytest.approx(l_frozen, rel=1e-6)
    gmax = max(float(sdg[k].grad.abs().max()) for k in sdg if k.startswith(hp) and sdg[k].grad is not None)
    for n, p in frozen.network.named_parameters():
        if n.startswith(hp):
            og = sdg[n].grad if sdg[n].grad is not None else torch.zeros_like(sd0[n])
            assert float((p.grad.cpu() - og).abs().max()) <= 5e-3 * gmax + 1e-9, n
            pf = dict(full.network.named_parameters())[n]
            assert float((p.grad - pf.grad).abs().max()) <= 1e-5 * gmax + 1e-12, n
        else:
            assert p.grad is None
    # switch to whole-network SGD
    frozen.epoch = frozen.warmup_duration
    frozen.on_epoch_end()
    assert isinstance(frozen.optimizer, torch.optim.SGD) and all(p.requires_grad for p in frozen.network.parameters())
    l0 = float(frozen.run_iteration(iter([batch]), True))
    losses = [float(frozen.run_iteration(iter([batch]), True)) for _ in range(6)]
    assert np.isfinite(losses).all() and losses[-1] < l0
    # softmax prediction through the residual network
    vol = np.random.RandomState(5).randn(1, 16, 40, 40).astype(np.float32)
    seg, prob = frozen.predict_preprocessed_data_return_seg_and_softmax(vol, do_mirroring=False, verbose=False)
    assert prob.shape == (4, 16, 40, 40) and seg.shape == (16, 40, 40)
    np.testing.assert_allclose(prob.sum(0), 1.0, atol=1e-4)
    assert frozen.network.decoder.deep_supervision is True
