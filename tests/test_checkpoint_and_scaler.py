"""Checkpoint interchange with the reference (network_trainer.py:256-286, nnUNetTrainerV2_DDP.py:636-697) and fp16 as the
reference runs it (dynamic loss scaling with GradScaler semantics, MultiTalent_Trainer_DDP.py:349-354)."""
import os
import pickle

import numpy as np
import pytest
import torch
from torch import nn

from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import FlatArena

REF_KEYS = {'epoch', 'state_dict', 'optimizer_state_dict', 'lr_scheduler_state_dict', 'plot_stuff', 'best_stuff'}


def _small_module():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv3d(1, 5, 3), nn.InstanceNorm3d(5, affine=True), nn.Conv3d(5, 3, 1, bias=False))


def test_arena_optimizer_state_is_torch_sgd_format():
    """The arena's momentum is written per parameter in torch.optim.SGD's state_dict format: the reference's
    `self.optimizer.load_state_dict(checkpoint['optimizer_state_dict'])` (network_trainer.py:377) consumes it, and the
    arena reads the reference's back."""
    m = _small_module()
    arena = FlatArena(m)
    arena.mom.copy_(torch.arange(arena.n, dtype=torch.float32) * 1e-3)
    arena.first = False
    sd = arena.sgd_state_dict(1e-2, 0.99, 3e-5)
    ref_opt = torch.optim.SGD(_small_module().parameters(), 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    ref_opt.load_state_dict(sd)                                    # what the reference does on resume
    bufs = [ref_opt.state[p]['momentum_buffer'] for p in ref_opt.param_groups[0]['params']]
    assert [tuple(b.shape) for b in bufs] == [tuple(p.shape) for p in m.parameters()]
    g = ref_opt.param_groups[0]
    assert g['momentum'] == 0.99 and g['nesterov'] is True and g['weight_decay'] == 3e-5
    # and back: a state_dict written by torch's SGD restores the arena's momentum (padding slots stay zero)
    arena2 = FlatArena(_small_module())
    arena2.load_sgd_state_dict(ref_opt.state_dict())
    assert not arena2.first
    assert torch.equal(arena2.mom, arena.mom * (arena2.mom != 0)) and float(arena2.mom.abs().sum()) > 0
    # before the first step there are no buffers: torch's format has an empty state
    fresh = FlatArena(_small_module())
    assert fresh.sgd_state_dict(1e-2, 0.99, 3e-5)['state'] == {}
    arena2.load_sgd_state_dict(fresh.sgd_state_dict(1e-2, 0.99, 3e-5))
    assert arena2.first and float(arena2.mom.abs().sum()) == 0.0


@pytest.mark.gpu
def test_checkpoint_round_trip_resumes_identically(tmp_path):
    """save -> load into a fresh trainer -> step == step of the uninterrupted trainer (weights, momentum, learning rate of
    the restored epoch, GradScaler state), and the file carries the reference's key set."""
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_batch
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    patch = (16, 32, 32)
    plans = default_plans(patch_size=patch, batch_size=2)
    plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4

    def make():
        t = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False)
        torch.manual_seed(0)
        t.initialize(True)
        return t
    a = make()
    batch = synthetic_batch(patch, 2, 0, a.deep_supervision_scales)
    data = torch.from_numpy(batch['data']).cuda()
    tgt = [torch.from_numpy(t).cuda() for t in batch['target']]
    valid = [p['valid_regions'] for p in batch['properties']]
    a.epoch = 7
    a.all_tr_losses = [1.0] * 8                       # the history the epoch loop keeps (epoch + 1 entries at save time)
    a.all_val_losses = [1.0] * 8
    a.maybe_update_lr(a.epoch)
    for _ in range(2):
        a.train_step(data, tgt, valid, True)
    f = str(tmp_path / "model_latest.model")
    a.save_checkpoint(f)
    ck = torch.load(f, map_location="cpu", weights_only=False)
    assert REF_KEYS <= set(ck) and ck['lr_scheduler_state_dict'] is None and ck['epoch'] == 8
    assert len(ck['plot_stuff']) == 4 and len(ck['best_stuff']) == 3
    assert set(ck['optimizer_state_dict']) == {'state', 'param_groups'}
    with open(f + ".pkl", "rb") as fh:
        info = pickle.load(fh)
    assert info['name'] == 'MultiTalent_trainer_ddp' and len(info['init']) == 11
    b = make()
    with torch.no_grad():
        for prm in b.network.parameters():             # make sure the load really restores everything
            prm.add_(0.5)
    b.load_checkpoint(f, train=True)
    assert b.epoch == 8 and b.lr == pytest.approx(1e-2 * (1 - 8 / 1000) ** 0.9)
    assert torch.equal(b.arena.flat, a.arena.flat) and torch.equal(b.arena.mom, a.arena.mom) and not b.arena.first
    a.epoch = 8
    a.maybe_update_lr(a.epoch)
    la = a.train_step(data, tgt, valid, True)[0]
    lb = b.train_step(data, tgt, valid, True)[0]
    assert float(la) == float(lb)
    assert float((a.arena.flat - b.arena.flat).abs().max()) < 1e-6
    # a DDP-prefixed checkpoint (module.*) loads too (nnUNetTrainerV2_DDP.py:645-661)
    ck['state_dict'] = {"module." + k: v for k, v in ck['state_dict'].items()}
    c = make()
    c.load_checkpoint_ram(ck, train=False)
    assert all(torch.equal(x.cpu(), y.cpu()) for x, y in zip(c.network.state_dict().values(),
                                                            torch.load(f, weights_only=False)['state_dict'].values()))
    ck['state_dict']['module.bogus'] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="unexpected keys"):
        make().load_checkpoint_ram(ck, train=False)


@pytest.mark.gpu
def test_fp16_dynamic_loss_scaling_skips_overflow_and_recovers():
    """fp16 storage + dynamic loss scale: an injected overflow (absurd scale) must skip the update and halve the scale
    exactly as torch's GradScaler does; the following clean steps update the parameters and agree with the fp32 step."""
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_batch
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import (DeviceGradScaler,
                                                                                    MultiTalent_trainer_ddp)
    patch = (16, 32, 32)
    plans = default_plans(patch_size=patch, batch_size=2)
    plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
    t16 = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False, fp16=True)      # the reference CLI default
    torch.manual_seed(0)
    t16.initialize(True)
    assert t16.native_dtype == torch.float16 and isinstance(t16.amp_grad_scaler, DeviceGradScaler)
    assert t16.amp_grad_scaler.get_scale() == 65536.0
    t32 = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False)
    torch.manual_seed(0)
    t32.initialize(True)
    assert t32.amp_grad_scaler is None
    batch = synthetic_batch(patch, 2, 0, t16.deep_supervision_scales)
    data = torch.from_numpy(batch['data']).cuda()
    tgt = [torch.from_numpy(t).cuda() for t in batch['target']]
    valid = [p['valid_regions'] for p in batch['properties']]
    before = t16.arena.flat.clone()
    t16.amp_grad_scaler.state[0] = 2.0 ** 60          # every fp16 gradient overflows
    t16.train_step(data, tgt, valid, True)
    st = t16.amp_grad_scaler.state.cpu()
    assert torch.equal(t16.arena.flat, before), "an overflowing step must not touch the parameters"
    assert float(st[0]) == 2.0 ** 59 and float(st[2]) == 1.0 and float(st[3]) == 1.0 and float(st[1]) == 0.0
    # back to a sane scale: the step happens, the scale stays, the growth tracker counts
    t16.amp_grad_scaler.state[0] = 65536.0
    l16 = t16.train_step(data, tgt, valid, True)[0]
    l32 = t32.train_step(data, tgt, valid, True)[0]
    st = t16.amp_grad_scaler.state.cpu()
    assert float(st[0]) == 65536.0 and float(st[1]) == 1.0 and float(st[2]) == 0.0
    assert abs(float(l16) - float(l32)) < 2e-2 * max(1.0, abs(float(l32)))
    upd16, upd32 = (t16.arena.flat - before), (t32.arena.flat - before)
    cos = float((upd16 * upd32).sum() / (upd16.norm() * upd32.norm()))
    assert cos > 0.98, "fp16 (scaled) update direction vs fp32 update: cosine %.4f" % cos
    assert 0.8 < float(upd16.norm() / upd32.norm()) < 1.25
    # growth after `growth_interval` clean steps; the state_dict is GradScaler's
    t16.amp_grad_scaler.growth_interval = 2
    t16.train_step(data, tgt, valid, True)
    assert t16.amp_grad_scaler.get_scale() == 131072.0
    torch.amp.GradScaler("cuda").load_state_dict(t16.amp_grad_scaler.state_dict())


@pytest.mark.gpu
def test_fp16_gradients_are_not_lost_to_underflow():
    """without scaling the CE gradient coefficient w0/nvox underflows fp16 at large patches.  With the dynamic
    scale the fp16 head-weight gradients track the fp32 ones."""
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_batch
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    patch = (64, 96, 96)      # 590 k voxels: 0.53 / nvox = 9e-7, i.e. fp16-subnormal territory without a loss scale
    res = {}
    for name, kw in (("fp32", {}), ("fp16", {"fp16": True})):
        t = MultiTalent_trainer_ddp(default_plans(patch_size=patch, batch_size=1), 0, 0, init_distributed=False, **kw)
        torch.manual_seed(0)
        t.initialize(True)
        t.lr = 0.0
        t.weight_decay = 0.0
        batch = synthetic_batch(patch, 1, 0, t.deep_supervision_scales)
        t.train_step(torch.from_numpy(batch['data']).cuda(), [torch.from_numpy(x).cuda() for x in batch['target']],
                     [p['valid_regions'] for p in batch['properties']], True)
        scale = t.amp_grad_scaler.get_scale() if t.amp_grad_scaler is not None else 1.0
        if name == "fp16":   # the clean step did not change the scale
            assert scale == 65536.0
        res[name] = (t.network.seg_outputs[-1].weight.grad.detach().float() / scale).cpu().numpy()
    a, b = res["fp32"].ravel(), res["fp16"].ravel()
    cos = float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert cos > 0.99 and 0.9 < np.linalg.norm(b) / np.linalg.norm(a) < 1.1, (cos, np.linalg.norm(b) / np.linalg.norm(a))
