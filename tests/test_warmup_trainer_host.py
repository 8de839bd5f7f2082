"""Host logic of the fine-tuning workflow against the reference's own functions: the
pretrained-weight loader and the three-phase learning-rate schedule of nnUNetTrainerV2_warmupsegheads.  No GPU."""
import os
import sys

import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multitalent_b200.run.load_pretrained_weights import load_pretrained_weights  # noqa: E402
from multitalent_b200.training.network_training.nnUNetTrainerV2_warmup import warmup_lr  # noqa: E402

HAVE_REF = os.path.isdir("/root/reference/nnunet")


class _Net(nn.Module):
    """Module with nnU-Net-style key names: a trunk (`conv_blocks_*`, `tu`) and heads (`seg_outputs`)."""

    def __init__(self, n_classes, width=4):
        super().__init__()
        self.conv_blocks_context = nn.ModuleList([nn.Conv3d(1, width, 3), nn.Conv3d(width, width, 3)])
        self.conv_blocks_localization = nn.ModuleList([nn.Conv3d(2 * width, width, 3)])
        self.tu = nn.ModuleList([nn.ConvTranspose3d(width, width, 2, 2, bias=False)])
        self.seg_outputs = nn.ModuleList([nn.Conv3d(width, n_classes, 1, bias=False)])


def _ckpt(net, prefix=""):
    return {'state_dict': {prefix + k: v.clone() for k, v in net.state_dict().items()}}


def test_loader_transfers_trunk_but_not_heads_and_strips_module_prefix():
    torch.manual_seed(0)
    src, dst = _Net(47), _Net(3)
    before = {k: v.clone() for k, v in dst.state_dict().items()}
    keys = load_pretrained_weights(dst, _ckpt(src, "module."))
    after = dst.state_dict()
    for k, v in src.state_dict().items():
        if k.startswith("seg_outputs."):
            assert k not in keys and torch.equal(after[k], before[k])      # another class count: never transferred
        else:
            assert k in keys and torch.equal(after[k], v)


def test_loader_rejects_an_incompatible_trunk_and_changes_nothing():
    torch.manual_seed(0)
    src, dst = _Net(47, width=6), _Net(3, width=4)
    before = {k: v.clone() for k, v in dst.state_dict().items()}
    with pytest.raises(RuntimeError, match="not compatible"):
        load_pretrained_weights(dst, _ckpt(src))
    for k, v in dst.state_dict().items():
        assert torch.equal(v, before[k])


def test_loader_keeps_parameter_storage():
    """Copies are in place: parameters that are views of a flat arena must stay views."""
    src, dst = _Net(47), _Net(47)
    ptrs = {k: v.data_ptr() for k, v in dst.state_dict().items()}
    load_pretrained_weights(dst, _ckpt(src))
    assert all(v.data_ptr() == ptrs[k] for k, v in dst.state_dict().items())
    assert all(torch.equal(a, b) for a, b in zip(src.state_dict().values(), dst.state_dict().values()))


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container)")
def test_loader_matches_the_reference_function(tmp_path):
    from oracle import ref_import
    ref_import.install()
    from nnunet.run.load_pretrained_weights import load_pretrained_weights as ref_load
    torch.manual_seed(1)
    src = _Net(47)
    a, b = _Net(3), _Net(3)
    b.load_state_dict(a.state_dict())
    f = str(tmp_path / "ckpt.model")
    torch.save(_ckpt(src, "module."), f)
    ref_load(a, f)
    load_pretrained_weights(b, f)
    for (k, x), y in zip(a.state_dict().items(), b.state_dict().values()):
        assert torch.equal(x, y), k


# known answers of nnUNetTrainerV2_warmupsegheads.maybe_update_lr (nnUNetTrainerV2_warmup.py:87-112) with its defaults
# warmup_duration 10, num_epochs_sgd_warmup 50, warmup_max_lr 5e-4, initial_lr 1e-2, max_num_epochs 1060
@pytest.mark.parametrize("epoch,lr", [
    (0, 5e-5), (4, 2.5e-4), (9, 5e-4),                       # heads only, linear to warmup_max_lr
    (10, 2e-4), (34, 5e-3), (59, 1e-2),                      # whole network, linear to initial_lr
    (60, 1e-2 * (1 - 1 / 1000) ** 0.9), (559, 1e-2 * (1 - 500 / 1000) ** 0.9), (1059, 0.0),   # poly over 1000 epochs
])
def test_warmup_lr_known_answers(epoch, lr):
    assert warmup_lr(epoch, 10, 50, 5e-4, 1e-2, 1060) == pytest.approx(lr, rel=1e-12, abs=1e-18)


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container)")
def test_warmup_lr_matches_the_reference_method():
    from oracle import ref_import
    ref_import.install()
    from nnunet.training.network_training.nnUNet_variants.pretraining.nnUNetTrainerV2_warmup import nnUNetTrainerV2_warmupsegheads as Ref  # noqa: E501

    class Stub:  # the attributes maybe_update_lr reads; no trainer construction (it would set up folders / data)
        warmup_duration, num_epochs_sgd_warmup, warmup_max_lr, initial_lr, max_num_epochs = 10, 50, 5e-4, 1e-2, 1060
        lr = None

        def __init__(self):
            self.optimizer = type("O", (), {"param_groups": [{"lr": None}]})()

        def print_to_log_file(self, *a, **k):
            pass
    for epoch in [0, 1, 9, 10, 11, 59, 60, 61, 200, 1058, 1059]:
        s = Stub()
        s.epoch = epoch
        Ref.maybe_update_lr(s)
        assert s.optimizer.param_groups[0]['lr'] == pytest.approx(warmup_lr(epoch, 10, 50, 5e-4, 1e-2, 1060), rel=1e-12)


def _drive_epochs(n_epochs):
    """Our trainer's epoch bookkeeping with a stub optimizer (no network): (lr, optimizer kind) each epoch trains with."""
    from multitalent_b200.training.network_training.nnUNetTrainerV2_warmup import nnUNetTrainerV2_warmupsegheads as Mine

    class T(Mine):
        def initialize_optimizer_and_scheduler(self, seg_heads_only=False):
            self.optimizer = type("O", (), {"param_groups": [{"lr": None}]})()
            self.kind = "adamw_heads" if seg_heads_only else "sgd_all"
    t = T(None, 0)
    t.initialize_optimizer_and_scheduler(True)
    t.maybe_update_lr()                     # nnUNetTrainerV2.run_training: maybe_update_lr(self.epoch) before epoch 0
    used = []
    for _ in range(n_epochs):
        used.append((t.optimizer.param_groups[0]['lr'], t.kind))
        t.on_epoch_end()
    return used


def test_epoch_loop_lr_sequence_known_answers():
    """network_trainer.py:482-490, 603-616: maybe_update_lr runs BEFORE the epoch counter is incremented, so epoch
    e >= 1 trains at warmup_lr(e - 1); epoch 10 is still heads-only AdamW at 5e-4, epoch 11 the first SGD epoch at 2e-4,
    and the last epoch (1059) still has a non-zero learning rate."""
    used = _drive_epochs(1060)
    f = lambda e: warmup_lr(e, 10, 50, 5e-4, 1e-2, 1060)  # noqa: E731
    assert used[0] == (pytest.approx(f(0)), "adamw_heads")
    for e in range(1, 1060):
        assert used[e][0] == pytest.approx(f(e - 1), rel=1e-12), e
        assert used[e][1] == ("adamw_heads" if e <= 10 else "sgd_all"), e
    assert used[10][0] == pytest.approx(5e-4) and used[11][0] == pytest.approx(2e-4)
    assert used[1059][0] > 0.0


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container)")
def test_epoch_loop_lr_sequence_matches_the_reference_call_order():
    """The reference's own methods driven in the reference's own order (run_training: maybe_update_lr(self.epoch) once,
    then per epoch on_epoch_end -> [optimizer switch at epoch == warmup_duration] -> maybe_update_lr() -> epoch += 1)."""
    from oracle import ref_import
    ref_import.install()
    from nnunet.training.network_training.nnUNet_variants.pretraining.nnUNetTrainerV2_warmup import nnUNetTrainerV2_warmupsegheads as Ref  # noqa: E501

    class Stub:
        warmup_duration, num_epochs_sgd_warmup, warmup_max_lr, initial_lr, max_num_epochs = 10, 50, 5e-4, 1e-2, 1060
        lr, epoch = None, 0

        def __init__(self):
            self.optimizer = type("O", (), {"param_groups": [{"lr": None}]})()

        def print_to_log_file(self, *a, **k):
            pass
    s = Stub()
    Ref.maybe_update_lr(s, 0)
    ref_used = []
    for _ in range(200):
        ref_used.append(s.optimizer.param_groups[0]['lr'])
        Ref.maybe_update_lr(s)      # NetworkTrainer.on_epoch_end (network_trainer.py:609), self.epoch not yet incremented
        s.epoch += 1                # network_trainer.py:490
    mine = _drive_epochs(200)
    for e, (a, b) in enumerate(zip(ref_used, mine)):
        assert a == pytest.approx(b[0], rel=1e-12), e
