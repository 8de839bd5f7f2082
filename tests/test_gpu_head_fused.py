"""mtb200_head_bwd_fused (loss pass 2 + head data gradient + head weight gradient in one kernel) against
(1) an independent torch fp32 evaluation of the same three steps on the same 16-bit operands and (2) the library's own
three separate passes, and -- through the trainer -- a whole training step with and without the fusion.
Reference semantics: generic_UNet.py:349-351 (heads), MultiTalent_Trainer_DDP.py:567-606 (loss)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(B, dims, Cin_p, dtype, windows, seed=0, accumulate=False):
    from multitalent_b200.dataset_conversion.Task100_MultiTalent import NUM_LABELS, region_bitmasks
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, H, W = dims
    nvox = D * H * W
    Cout, Cout_p, C8 = 47, 48, 48
    dev = "cuda"
    z = (2.0 * torch.randn(B, nvox, Cout_p, generator=g)).to(dtype).to(dev)
    x = torch.randn(B, nvox, Cin_p, generator=g).to(dtype).to(dev)
    w = (0.1 * torch.randn(Cout, Cin_p, generator=g))
    w_swap = torch.zeros(Cin_p, Cout_p)
    w_swap[:, :Cout] = w.t()
    w_swap = w_swap.to(dtype).to(dev)
    tgt = torch.randint(0, NUM_LABELS, (B, nvox), generator=g).float().to(dev)
    pos, _ = region_bitmasks()
    pos_t = torch.tensor(pos, dtype=torch.int64, device=dev)
    coef = torch.zeros(B, C8, 4)
    for b, (c0, chans) in enumerate(windows):
        for j in chans:
            coef[b, j] = torch.tensor([1.0 / nvox, 0.3 / nvox * (1 + j % 3), 0.1 / nvox, 1.0])
    coef = coef.to(dev)
    gs = torch.tensor(512.0, device=dev)
    dx0 = torch.randn(B, nvox, Cin_p, generator=g).to(dtype).to(dev) if accumulate else None
    return dict(z=z, x=x, w_swap=w_swap, tgt=tgt, pos=pos_t, pos_list=pos, coef=coef, gs=gs, dx0=dx0, nvox=nvox,
                Cout_p=Cout_p, C8=C8, n_labels=NUM_LABELS)


def _reference(pr, dtype, windows):
    """torch fp32 on the same operands: d(logits) rounded to `dtype`, then the two GEMMs."""
    z, x, coef = pr["z"].float(), pr["x"].float(), pr["coef"]
    B, nvox, Cp = z.shape
    lab = pr["tgt"].long()
    pos = pr["pos"]
    y = ((pos[lab].unsqueeze(-1) >> torch.arange(Cp, device=z.device)) & 1).float()  # [B, nvox, Cp]
    sig = torch.sigmoid(z)
    c = coef * torch.tensor([1.0, 1.0, 1.0, 0.0], device=z.device) * pr["gs"]
    d = c[:, None, :, 0] * (sig - y) - sig * (1 - sig) * (y * c[:, None, :, 1] - c[:, None, :, 2])
    d = d * (coef[:, None, :, 3] != 0)
    d = d.to(dtype).float()
    wsw = pr["w_swap"].float()              # [Cin, Cout_p]
    dx = d @ wsw.t()                        # [B, nvox, Cin]
    if pr["dx0"] is not None:
        dx = dx.to(dtype).float() + pr["dx0"].float()
    dw = torch.einsum("bvj,bvc->jc", d, x)  # [Cout_p, Cin]
    return d, dx, dw


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("Cin_p,dims,accumulate", [(32, (8, 16, 24), False), (64, (6, 10, 16), True),
                                                   (32, (5, 7, 9), True)])
def test_head_bwd_fused_matches_torch(dtype, Cin_p, dims, accumulate):
    from multitalent_b200 import _lib as L
    B = 3
    windows = [(0, [0, 1]), (8, list(range(9, 22))), (32, [43, 44, 45, 46])]
    pr = _problem(B, dims, Cin_p, dtype, windows, seed=Cin_p + dims[0], accumulate=accumulate)
    dx = pr["dx0"].clone() if accumulate else torch.full_like(pr["x"], float("nan"))
    dw = torch.zeros(pr["Cout_p"], Cin_p, device="cuda")
    p = L.HeadBwdParams()
    p.logits, p.target, p.coef, p.gscale = pr["z"].data_ptr(), pr["tgt"].data_ptr(), pr["coef"].data_ptr(), pr["gs"].data_ptr()
    p.pos_mask, p.x, p.w_swap, p.dx, p.dw = (pr["pos"].data_ptr(), pr["x"].data_ptr(), pr["w_swap"].data_ptr(),
                                             dx.data_ptr(), dw.data_ptr())
    p.nvox, p.dtype, p.B = pr["nvox"], L.dtype_enum(dtype), B
    p.z_ldc, p.C8, p.n_labels = pr["Cout_p"], pr["C8"], pr["n_labels"]
    p.x_ldc, p.x_coff, p.Cin, p.Cout = Cin_p, 0, Cin_p, pr["Cout_p"]
    p.dx_ldc, p.dx_coff, p.accumulate = Cin_p, 0, int(accumulate)
    for b, (c0, _) in enumerate(windows):
        p.win_c0[b] = c0
    L.call("mtb200_head_bwd_fused", C.byref(p), L.stream_ptr())
    torch.cuda.synchronize()
    assert L.lib().mtb200_last_kernel() == b"head_bwd_fused"
    d, dx_ref, dw_ref = _reference(pr, dtype, windows)
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    tol = 2.5 * ulp * float(dx_ref.abs().max())
    assert torch.isfinite(dx.float()).all()
    err = float((dx.float() - dx_ref).abs().max())
    assert err <= tol, "dx: %.3e > %.3e" % (err, tol)
    errw = float((dw - dw_ref).abs().max())
    assert errw <= 2e-3 * float(dw_ref.abs().max()), "dw: %.3e vs max %.3e" % (errw, float(dw_ref.abs().max()))
    # nothing outside the windows
    outside = torch.ones(pr["Cout_p"], dtype=torch.bool)
    for c0, chans in windows:
        outside[chans] = False
    assert float(dw[outside.cuda()].abs().max()) == 0.0


def test_head_bwd_fused_matches_separate_passes():
    """The library's own three passes (mtb200_mt_loss_bwd, pointwise data gradient, head weight gradient) on the same
    problem: d(input) to 1 ulp of bf16, d(weight) to fp32 summation order."""
    from multitalent_b200 import _lib as L
    from multitalent_b200.engine import ConvOp, Engine, Feat, Tape
    from torch import nn
    dtype, Cin_p, dims, B = torch.bfloat16, 32, (8, 16, 16), 2
    windows = [(8, list(range(9, 22))), (24, [30, 31, 32, 33, 34, 35])]
    pr = _problem(B, dims, Cin_p, dtype, windows, seed=3)
    nvox = pr["nvox"]
    conv = nn.Conv3d(30, 47, 1, bias=False).cuda()
    op = ConvOp(conv.weight, None, (1, 1, 1), (1, 1, 1))
    eng = Engine(dtype, 0)
    w_swap = op.packed(dtype, True)[0]  # [Cin_p][Cout_p]
    # fused
    dx = torch.empty_like(pr["x"])
    dw = torch.zeros(48, Cin_p, device="cuda")
    p = L.HeadBwdParams()
    p.logits, p.target, p.coef, p.gscale = pr["z"].data_ptr(), pr["tgt"].data_ptr(), pr["coef"].data_ptr(), pr["gs"].data_ptr()
    p.pos_mask, p.x, p.w_swap, p.dx, p.dw = pr["pos"].data_ptr(), pr["x"].data_ptr(), w_swap.data_ptr(), dx.data_ptr(), dw.data_ptr()
    p.nvox, p.dtype, p.B = nvox, L.dtype_enum(dtype), B
    p.z_ldc, p.C8, p.n_labels = 48, 48, pr["n_labels"]
    p.x_ldc, p.x_coff, p.Cin, p.Cout = Cin_p, 0, Cin_p, 48
    p.dx_ldc, p.dx_coff, p.accumulate = Cin_p, 0, 0
    for b, (c0, _) in enumerate(windows):
        p.win_c0[b] = c0
    L.call("mtb200_head_bwd_fused", C.byref(p), L.stream_ptr())
    # separate passes
    D, H, W = dims
    dz = torch.empty(B, D, H, W, 48, dtype=dtype, device="cuda")
    L.call("mtb200_mt_loss_bwd", pr["z"].data_ptr(), L.dtype_enum(dtype), 48, 48, pr["tgt"].data_ptr(), B, nvox,
           pr["pos"].data_ptr(), pr["n_labels"], pr["coef"].data_ptr(), pr["gs"].data_ptr(), dz.data_ptr(), 48,
           L.stream_ptr())
    x = Feat(pr["x"].view(B, D, H, W, Cin_p), 0, 30, Cin_p)
    dyf = Feat(dz, 0, 47, 48)
    tape = Tape()
    tape.grad_bufs[id(dz)] = dz
    tape.grad_init[id(dz)] = set()
    gx, _ = tape.grad_feat(x)
    eng._conv_call(op.dgrad_taps, dyf, op.packed(dtype, True), None, gx, gx.dims[1:], None, False, op.Cout_p, op.Cin_p)
    dw2 = torch.zeros(1, 48, Cin_p, device="cuda")
    eng.overlap_wgrad = False
    eng._wgrad(tape, op, x, dyf, dw2, L.dtype_enum(dtype), x.buf.device, True, True)
    torch.cuda.synchronize()
    ref = gx.buf.view(B, nvox, Cin_p).float()
    err = float((dx.float() - ref).abs().max())
    assert err <= 2.0 ** -8 * float(ref.abs().max()), err
    errw = float((dw - dw2[0]).abs().max())
    assert errw <= 1e-4 * float(dw2.abs().max()), (errw, float(dw2.abs().max()))


def test_training_step_with_and_without_head_fusion():
    """Same state, same batch: the step with the fused head backward leaves the same loss and (to 16-bit rounding noise)
    the same gradients in the arena as the step with the three separate passes."""
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_batch
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    patch, B = (16, 32, 32), 3
    plans = default_plans(patch_size=patch, batch_size=B)
    plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
    res = {}
    for fuse in (False, True):
        tr = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False, native_dtype=torch.bfloat16)
        torch.manual_seed(0)
        tr.initialize(True)
        tr.lr = 0.0
        tr.weight_decay = 0.0
        eng = tr.network._engine
        eng.fuse_head = fuse
        batch = synthetic_batch(patch, B, 6, tr.deep_supervision_scales)  # rank 6: datasets 5, 6, 7 (1, 13 and 8 regions)
        data = torch.from_numpy(batch['data']).cuda()
        tgt = [torch.from_numpy(t).cuda() for t in batch['target']]
        valid = [p['valid_regions'] for p in batch['properties']]
        from multitalent_b200 import _lib as L
        with L.KernelProfile() as kp:
            l, ce, dc = tr.train_step(data, tgt, valid, True)
        names = {r[5] for r in kp.per_launch_kernels()}
        assert ("head_bwd_fused" in names) == fuse, names
        res[fuse] = (float(l), tr.arena.grad.clone(), [(n, p.grad.detach().clone()) for n, p in tr.network.named_parameters()])
    assert res[True][0] == res[False][0]
    ga, gb = res[False][1].double(), res[True][1].double()
    cos = float((ga * gb).sum() / (ga.norm() * gb.norm()))
    assert cos > 0.99999, cos
    for (n, a), (_, b) in zip(res[False][2], res[True][2]):
        scale = float(a.abs().max())
        if scale < 1e-6 * float(ga.abs().max()):
            continue
        assert float((a - b).abs().max()) <= 2e-2 * scale, (n, float((a - b).abs().max()), scale)


@pytest.mark.parametrize("Cin_p,dims,hard", [(32, (8, 16, 24), True), (64, (6, 10, 16), False), (32, (5, 7, 9), False)])
def test_head_fwd_stats_matches_separate_passes(Cin_p, dims, hard):
    """mtb200_head_fwd_stats (logits window on the tensor core, never stored) against pointwise head + mtb200_mt_loss_stats
    on the stored 16-bit logits, and against a torch evaluation of the same sums."""
    from multitalent_b200 import _lib as L
    dtype, B = torch.bfloat16, 3
    windows = [(0, [0, 1]), (8, list(range(9, 22))), (32, [43, 44, 45, 46])]
    pr = _problem(B, dims, Cin_p, dtype, windows, seed=7 + Cin_p)
    nvox = pr["nvox"]
    w_fwd = pr["w_swap"].t().contiguous()          # [Cout_p][Cin]
    valid = torch.tensor([sum(1 << j for j in chans) for _, chans in windows], dtype=torch.int64, device="cuda")
    stats = torch.zeros(B, 48, 4, dtype=torch.float64, device="cuda")
    hd = torch.zeros(B, 48, 2, dtype=torch.float64, device="cuda") if hard else None
    p = L.HeadFwdParams()
    p.x, p.w_fwd, p.target = pr["x"].data_ptr(), w_fwd.data_ptr(), pr["tgt"].data_ptr()
    p.valid_mask, p.pos_mask, p.stats = valid.data_ptr(), pr["pos"].data_ptr(), stats.data_ptr()
    p.hard = hd.data_ptr() if hard else None
    p.nvox, p.dtype, p.B, p.C8, p.n_labels = nvox, L.dtype_enum(dtype), B, 48, pr["n_labels"]
    p.x_ldc, p.x_coff, p.Cin, p.Cout = Cin_p, 0, Cin_p, 48
    for b, (c0, _) in enumerate(windows):
        p.win_c0[b] = c0
    L.call("mtb200_head_fwd_stats", C.byref(p), L.stream_ptr())
    # separate passes: logits = x @ W^T rounded to bf16 (what the pointwise kernel stores), then the statistics kernel
    z = (pr["x"].float() @ w_fwd.float().t()).to(dtype).contiguous()
    stats2 = torch.zeros_like(stats)
    hd2 = torch.zeros(B, 48, 2, dtype=torch.float64, device="cuda") if hard else None
    L.call("mtb200_mt_loss_stats", z.data_ptr(), L.dtype_enum(dtype), 48, 48, pr["tgt"].data_ptr(), B, nvox,
           valid.data_ptr(), pr["pos"].data_ptr(), pr["n_labels"], stats2.data_ptr(),
           hd2.data_ptr() if hard else None, L.stream_ptr())
    torch.cuda.synchronize()
    scale = stats2.abs().amax(dim=(0, 1)).clamp_min(1e-9)
    err = ((stats - stats2).abs() / scale).max()
    assert float(err) < 2e-4, (float(err), stats[0, :2], stats2[0, :2])
    if hard:
        # a logit within rounding of 0 may flip its thresholded prediction between the two accumulation orders
        assert float((hd - hd2).abs().max()) <= 2.0, (hd - hd2).abs().max()
    # torch evaluation of sum sigma * y and sum y on the stored logits
    lab = pr["tgt"].long()
    y = ((pr["pos"][lab].unsqueeze(-1) >> torch.arange(48, device="cuda")) & 1).double()
    sig = torch.sigmoid(z.double())
    for b, (c0, chans) in enumerate(windows):
        for j in chans:
            assert abs(float(stats[b, j, 1]) - float((sig[b, :, j] * y[b, :, j]).sum())) <= 1e-3 * max(1.0, float(y[b, :, j].sum()))
            assert float(stats[b, j, 3]) == float(y[b, :, j].sum())


def test_training_step_with_deferred_heads():
    """Heads deferred into the loss (no logits in memory) vs heads computed by the network: same loss, same gradients."""
    from multitalent_b200 import _lib as L
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_batch
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    patch, B = (16, 32, 32), 3
    plans = default_plans(patch_size=patch, batch_size=B)
    plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
    res = {}
    for defer in (False, True):
        tr = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False, native_dtype=torch.bfloat16)
        torch.manual_seed(0)
        tr.initialize(True)
        tr.lr = 0.0
        tr.weight_decay = 0.0
        tr.network._engine.defer_head_fwd = defer
        batch = synthetic_batch(patch, B, 6, tr.deep_supervision_scales)
        data = torch.from_numpy(batch['data']).cuda()
        tgt = [torch.from_numpy(t).cuda() for t in batch['target']]
        valid = [p['valid_regions'] for p in batch['properties']]
        with L.KernelProfile() as kp:
            l, ce, dc = tr.train_step(data, tgt, valid, True)
        names = {r[5] for r in kp.per_launch_kernels()}
        assert ("head_fwd_stats" in names) == defer, names
        res[defer] = (float(l.detach()), float(ce), float(dc), tr.arena.grad.clone())
        # a step that keeps its output (online evaluation) must not defer
        l2, _, _ = tr.train_step(data, tgt, valid, True, keep_output=True)
        assert tr._last_output[0].shape[1] == 47 and float(tr._last_output[0].float().abs().max()) > 0
    for k in range(3):
        assert abs(res[True][k] - res[False][k]) <= 1e-5 * max(1.0, abs(res[False][k])), (k, res[True][k], res[False][k])
    ga, gb = res[False][3].double(), res[True][3].double()
    cos = float((ga * gb).sum() / (ga.norm() * gb.norm()))
    assert cos > 0.99999, cos


@pytest.mark.parametrize("mirror", [False, True])
def test_sliding_window_with_fused_head_aggregate(mirror):
    """predict_3D with head -> sigmoid x Gaussian -> scatter-add as one kernel per tile (mtb200_head_aggregate) against the
    same predictor with the pointwise head + mtb200_sw_aggregate: the accumulated probabilities agree to fp32 rounding."""
    from multitalent_b200 import _lib as L
    from multitalent_b200.plans import default_plans
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    patch = (16, 32, 32)
    plans = default_plans(patch_size=patch, batch_size=2)
    plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
    tr = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False, native_dtype=torch.bfloat16)
    torch.manual_seed(0)
    tr.initialize(False)
    net = tr.network
    net.eval()
    net.do_ds = False
    vol = np.random.RandomState(3).randn(1, 24, 40, 56).astype(np.float32)
    kw = dict(do_mirroring=mirror, mirror_axes=(0, 1, 2), use_sliding_window=True, step_size=0.5, patch_size=patch,
              regions_class_order=tuple(range(47)), use_gaussian=True, verbose=False, return_device_tensors=True)
    res = {}
    for fuse in (False, True):
        net._engine.fuse_head_aggregate = fuse
        with L.KernelProfile() as kp:
            seg, prob = net.predict_3D(vol, **kw)
        names = {r[5] for r in kp.per_launch_kernels()}
        assert ("head_aggregate" in names) == fuse, names
        res[fuse] = (seg.clone(), prob.clone())
    assert float((res[True][1] - res[False][1]).abs().max()) <= 2e-6
    assert torch.equal(res[True][0], res[False][0])
