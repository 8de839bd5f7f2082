"""Parity checks at BASELINE.json's FULL sizes through size-independent properties (the CPU oracle cannot run these sizes
in seconds): (1) the tensor-core (bf16) training step against the CUDA-core fp32 step of the same network on the same
192x160x128 batch -- every full-resolution kernel family (first-layer, line-streaming, group-merged, pointwise, per-tap,
line / per-tap weight gradients) at its real shapes; (2) partition of unity of the sliding-window aggregation: a network
whose logits are constant per class must give exactly sigmoid(constant) everywhere, for every overlap pattern, with and
without the 8 mirror passes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
FULL_PATCH = (192, 160, 128)


def _trainer(dtype, batch):
    from multitalent_b200.plans import default_plans
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    tr = MultiTalent_trainer_ddp(default_plans(patch_size=FULL_PATCH, batch_size=batch), 0, 0, native_dtype=dtype,
                                 init_distributed=False)
    torch.manual_seed(0)
    tr.initialize(True)
    tr.lr = 0.0  # keep the parameters: the step only has to leave its gradients in the arena
    tr.weight_decay = 0.0
    return tr


def test_full_patch_bf16_step_matches_fp32_step():
    from multitalent_b200.synthetic import synthetic_batch
    B = 2
    t32, t16 = _trainer(torch.float32, B), _trainer(torch.bfloat16, B)
    assert torch.equal(t32.arena.flat, t16.arena.flat), "same seed must give the same initial weights"
    batch = synthetic_batch(FULL_PATCH, B, 0, t32.deep_supervision_scales)
    data = torch.from_numpy(batch['data']).cuda()
    tgt = [torch.from_numpy(t).cuda() for t in batch['target']]
    valid = [p['valid_regions'] for p in batch['properties']]
    res = {}
    for name, tr in (("fp32", t32), ("bf16", t16)):
        l, ce, dc = tr.train_step(data, tgt, valid, True)
        torch.cuda.synchronize()
        res[name] = (float(l), float(ce), float(dc), tr.arena.grad.clone(),
                     [(n, p.grad.detach().clone()) for n, p in tr.network.named_parameters()])
        del tr
    l32, l16 = res["fp32"], res["bf16"]
    assert np.isfinite(l16[0]) and abs(l16[0] - l32[0]) <= 1e-2 * max(1.0, abs(l32[0])), (l16[:3], l32[:3])
    assert abs(l16[1] - l32[1]) <= 1e-2 * max(1.0, abs(l32[1])) and abs(l16[2] - l32[2]) <= 1e-2 * max(1.0, abs(l32[2]))
    g32, g16 = l32[3].double(), l16[3].double()
    cos = float((g32 * g16).sum() / (g32.norm() * g16.norm()))
    assert cos > 0.98, "gradient cosine (all parameters) bf16 vs fp32 = %.4f" % cos
    worst = 1.0
    for (n, a), (_, b) in zip(l32[4], l16[4]):
        a, b = a.double().flatten(), b.double().flatten()
        if float(a.norm()) < 1e-6 * float(g32.norm()):
            continue  # biases in front of an InstanceNorm (exact zeros) and other negligible tensors
        c = float((a * b).sum() / (a.norm() * b.norm() + 1e-300))
        worst = min(worst, c)
        # measured: 0.915 for the worst tensor (a deep, few-voxel layer), 0.9986 over all parameters; the forward pass is
        # reproducible, so these numbers do not move from run to run
        assert c > 0.85, "gradient cosine of %s = %.4f" % (n, c)
    print("full-patch bf16 vs fp32: loss %.6f vs %.6f, cosine %.5f, worst tensor %.4f" % (l16[0], l32[0], cos, worst))


@pytest.mark.parametrize("mirror", [False, True])
def test_sliding_window_partition_of_unity_full_size(mirror):
    from multitalent_b200.engine import Feat
    from multitalent_b200.plans import default_plans
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    tr = MultiTalent_trainer_ddp(default_plans(patch_size=FULL_PATCH, batch_size=1), 0, 0, native_dtype=torch.bfloat16,
                                 init_distributed=False)
    tr.initialize(False)
    net = tr.network
    net.eval()
    net.do_ds = False
    consts = torch.linspace(-3.0, 3.0, 47)
    consts_b = consts.to(torch.bfloat16).float()  # what the logits buffer really holds

    def const_logits(tile):
        B, D, H, W = tile.dims
        buf = torch.zeros((B, D, H, W, 48), dtype=torch.bfloat16, device=tile.buf.device)
        buf[..., :47] = consts.to(tile.buf.device).to(torch.bfloat16)
        return Feat(buf, 0, 47, 48)
    net.native_logits = const_logits
    vol = np.zeros((1, 300, 350, 270), dtype=np.float32)  # 3 x 4 x 4 tiles, ragged overlaps in every axis
    seg, prob = net.predict_3D(vol, do_mirroring=mirror, mirror_axes=(0, 1, 2), use_sliding_window=True, step_size=0.5,
                               patch_size=FULL_PATCH, regions_class_order=tuple(range(47)), use_gaussian=True,
                               verbose=False, return_device_tensors=True)
    want = torch.sigmoid(consts_b).to(prob.device)
    err = (prob - want.view(47, 1, 1, 1)).abs().amax(dim=(1, 2, 3))
    assert float(err.max()) < 5e-6, "aggregated probabilities deviate from sigmoid(const) by %.3e" % float(err.max())
    last_on = max(j for j in range(47) if float(want[j]) > 0.5)
    assert torch.all(seg == float(last_on)), "in-order threshold (neural_network.py:415-417) must leave class %d" % last_on
