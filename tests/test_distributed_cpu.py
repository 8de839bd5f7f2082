"""CPU / gloo world-size-2 tests of the N>1 host logic: the packed all-gather pooling of the Dice statistics and the
closed-form (x world size) gradient rule, checked against the UNMODIFIED reference running under real 2-rank gloo when
/root/reference is present, and against the oracle's single-process emulation everywhere."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_inputs(rank):
    from oracle import unet_oracle as O
    g = torch.Generator().manual_seed(100 + rank)
    shapes = [(2, 47, 4, 8, 8), (2, 47, 2, 4, 4), (2, 47, 1, 2, 2)]
    logits = [torch.randn(s, generator=g) * 2 for s in shapes]
    tasks = [O.TASK_IDS[(2 * rank + b) % 13] for b in range(2)]
    if rank == 1:
        tasks[0] = O.TASK_IDS[0]  # make both ranks supervise the liver channels at b=0 -> pooled Dice really mixes ranks
    if rank == 0:
        tasks[0] = O.TASK_IDS[0]
    rng = np.random.RandomState(7 + rank)
    targets = []
    for s in shapes:
        t = np.zeros((2, 1) + s[2:], dtype=np.float32)
        for b in range(2):
            labs = O.TASK_LABEL_MAPS[tasks[b]][1]
            t[b, 0] = rng.choice([0] + list(labs), size=s[2:])
        targets.append(torch.from_numpy(t))
    valid = [O.VALID_REGIONS[t] for t in tasks]
    return logits, targets, valid


def _local_packed(logits, targets, valid):
    """[scales, B, 47, 2] = {tp, sum sigma + sum y} -- what the native loss packs for the all-gather."""
    from oracle import unet_oracle as O
    out = []
    for z, t in zip(logits, targets):
        _, tp, fp, fn = O.multitalent_loss_stats(z, t, valid)
        out.append(torch.stack((tp, 2 * tp + fp + fn), -1))
    return torch.stack(out, 0).double()


def _worker(rank, world, initfile, use_reference, q):
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=initfile, rank=rank, world_size=world)
    try:
        from oracle import unet_oracle as O
        from multitalent_b200.training.loss_functions.multitalent_loss import pool_stats_over_ranks
        logits, targets, valid = _rank_inputs(rank)
        w = O.multitalent_ds_loss_weights(3)
        packed = _local_packed([z.detach() for z in logits], targets, valid)
        pooled = pool_stats_over_ranks(packed)          # product host logic under a real 2-rank gloo group
        other = (pooled - packed).float()               # statistics of all OTHER ranks
        # oracle emulation of this rank's loss + gradient (Dice gradient x world size)
        zs = [z.clone().requires_grad_(True) for z in logits]
        other_stats = []
        for i in range(3):
            otp = other[i, ..., 0]
            od = other[i, ..., 1]
            # split D_other = 2tp+fp+fn back into (tp, fp, fn)-compatible pieces: only tp and the sum matter
            other_stats.append((otp, od - 2 * otp, torch.zeros_like(otp)))
        l, ce, dc = O.multitalent_loss_ddp(zs, targets, valid, w, other_stats, world)
        l.backward()
        res = {"loss": l.item(), "ce": ce.item(), "dc": dc.item(), "grads": [z.grad.numpy().copy() for z in zs],
               "pooled": pooled.numpy().copy()}  # numpy: pickled by value (torch tensors travel via /dev/shm handles)
        if use_reference:
            from oracle import ref_import
            ref_import.install()
            zr = [z.clone().requires_grad_(True) for z in logits]
            lr_, cer, dcr = ref_import.reference_compute_loss(zr, targets, valid, w)
            lr_.backward()
            res["ref"] = {"loss": lr_.item(), "ce": cer.item(), "dc": dcr.item(), "grads": [z.grad.numpy().copy() for z in zr]}
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _run(world, use_reference):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    init = "tcp://127.0.0.1:%d" % port
    procs = [ctx.Process(target=_worker, args=(r, world, init, use_reference, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r, res = q.get(timeout=240)
        out[r] = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_pooling_sums_over_ranks_per_batch_index():
    out = _run(2, False)
    sys.path.insert(0, ROOT)
    packed = [_local_packed(*_rank_inputs(r)) for r in range(2)]
    expect = packed[0] + packed[1]                      # (W, B, C).sum(0): pooled over ranks, NOT over the batch
    for r in range(2):
        np.testing.assert_allclose(out[r]["pooled"], expect.numpy(), rtol=1e-12)
    # both ranks share the pooled Dice but have their own CE
    assert abs(out[0]["dc"] - out[1]["dc"]) < 1e-5
    assert abs(out[0]["ce"] - out[1]["ce"]) > 1e-3


@pytest.mark.skipif(not os.path.isdir("/root/reference/nnunet"), reason="reference only in the build container")
def test_two_rank_loss_and_gradient_match_reference_under_gloo():
    out = _run(2, True)
    for r in range(2):
        res, ref = out[r], out[r]["ref"]
        assert abs(res["loss"] - ref["loss"]) < 1e-4 and abs(res["dc"] - ref["dc"]) < 1e-4
        for g, gr in zip(res["grads"], ref["grads"]):
            scale = float(np.abs(gr).max()) + 1e-12
            assert float(np.abs(g - gr).max()) / scale < 1e-4


# ---- online evaluation under two ranks (MT:372-410: per-rank counts all-gathered to [W, B, 47]) ---------------------------
def _eval_worker(rank, world, initfile, use_reference, q):
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=initfile, rank=rank, world_size=world)
    try:
        from types import SimpleNamespace
        from multitalent_b200.training.online_evaluation import OnlineEvaluationMixin
        logits, targets, valid = _rank_inputs(rank)

        class Mine(OnlineEvaluationMixin):
            pass
        mine = Mine()
        mine.run_online_evaluation([logits[0]], [targets[0]], valid)
        res = {k: np.array(getattr(mine, k), dtype=np.float64) for k in
               ("online_eval_foreground_dc", "online_eval_tp", "online_eval_fp", "online_eval_fn")}
        if use_reference:
            from oracle import ref_import
            ref_import.install()
            from nnunet.training.network_training.custom_trainers.MultiTalent.MultiTalent.MultiTalent_Trainer_DDP import \
                MultiTalent_trainer_ddp as Ref
            ref = SimpleNamespace(online_eval_foreground_dc=[], online_eval_tp=[], online_eval_fp=[], online_eval_fn=[])
            Ref.run_online_evaluation(ref, [logits[0]], [targets[0]], valid)
            res["ref"] = {k: np.array(getattr(ref, k), dtype=np.float64) for k in
                          ("online_eval_foreground_dc", "online_eval_tp", "online_eval_fp", "online_eval_fn")}
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _run_eval(world, use_reference):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    init = "tcp://127.0.0.1:%d" % port
    procs = [ctx.Process(target=_eval_worker, args=(r, world, init, use_reference, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r, res = q.get(timeout=240)
        out[r] = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_online_evaluation_gathers_counts_over_ranks():
    sys.path.insert(0, ROOT)
    from multitalent_b200.training.online_evaluation import hard_tp_fp_fn
    out = _run_eval(2, os.path.isdir("/root/reference/nnunet"))
    local = []
    for r in range(2):
        logits, targets, valid = _rank_inputs(r)
        local.append([t.numpy().astype(np.float64) for t in hard_tp_fp_fn(logits[0], targets[0], valid)])
    tp = np.stack([local[0][0], local[1][0]])            # [W, B, 47]
    for r in range(2):
        assert out[r]["online_eval_foreground_dc"].shape == (1, 2, 2, 47)   # one iteration, [W, B, 47]
        np.testing.assert_allclose(out[r]["online_eval_tp"][0], tp.sum(0), rtol=0, atol=0)
        if "ref" in out[r]:
            for k, v in out[r]["ref"].items():
                np.testing.assert_allclose(out[r][k], v, rtol=1e-6, atol=1e-7, err_msg=k)
