"""Two-GPU data-parallel correctness (needs 2 CUDA devices; skipped otherwise): one process per GPU over NCCL, per-rank
batches.  (1) after 3 steps the parameters are bit-identical on both ranks; (2) the exchanged gradient equals the mean of
the two ranks' local gradients (what torch DDP's reducer computes, MultiTalent_Trainer_DDP.py:121) -- with the pooled-Dice
semantics of the loss (statistics all-gathered, Dice gradient x world size) included; (3) the overlapped two-part
all-reduce gives the same parameters as one all-reduce after the backward pass."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, early, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), MTB200_EARLY_ALLREDUCE="1" if early else "0")
    import torch.distributed as dist
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_batch
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp
    patch = (32, 64, 64)
    tr = MultiTalent_trainer_ddp(default_plans(patch_size=patch, batch_size=2), 0, rank, native_dtype=torch.bfloat16)
    torch.manual_seed(100 + rank)          # different initial weights per rank: the construction broadcast must fix that
    tr.initialize(True)
    assert tr.world_size == world and (tr._early_ranges is not None) == early
    b = synthetic_batch(patch, 2, rank, tr.deep_supervision_scales)
    x = torch.from_numpy(b['data']).cuda()
    tg = [torch.from_numpy(t).cuda() for t in b['target']]
    valid = [p['valid_regions'] for p in b['properties']]
    # local gradient of this rank (no exchange): run forward/backward by hand on the same state
    tr.arena.zero_grad()
    out = tr.network(x)
    l, _, _ = tr.compute_loss(out, tg, valid)
    l.backward()
    torch.cuda.synchronize()
    if early:  # the backward pass already exchanged the early ranges: undo nothing, just finish -> arena.grad = SUM over ranks
        tr._finish_allreduce()
        summed = tr.arena.grad.clone()
        local = None
    else:
        local = tr.arena.grad.clone()
        dist.all_reduce(tr.arena.grad)
        summed = tr.arena.grad.clone()
    losses = []
    for _ in range(3):
        losses.append(float(tr.train_step(x, tg, valid, True)[0]))
    torch.cuda.synchronize()
    torch.save({"flat": tr.arena.flat.cpu(), "summed": summed.cpu(), "local": None if local is None else local.cpu(),
                "losses": losses}, os.path.join(out_dir, "r%d_e%d.pt" % (rank, int(early))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_training_matches_across_ranks_and_exchange_modes(tmp_path):
    import torch.multiprocessing as mp
    res = {}
    for early in (False, True):
        mp.spawn(_worker, args=(2, _free_port(), early, str(tmp_path)), nprocs=2, join=True)
        res[early] = [torch.load(str(tmp_path / ("r%d_e%d.pt" % (r, int(early))))) for r in range(2)]
    for early in (False, True):
        a, b = res[early]
        assert torch.equal(a["flat"], b["flat"]), "parameters differ between the ranks after 3 steps (early=%s)" % early
        assert torch.equal(a["summed"], b["summed"])
        assert np.isfinite(a["losses"]).all() and a["losses"][-1] < a["losses"][0]
    # exchanged gradient = sum of the two local gradients (the SGD kernel divides by the world size)
    a, b = res[False]
    want = a["local"] + b["local"]
    assert float((a["summed"] - want).abs().max()) <= 1e-6 * float(want.abs().max()) + 1e-12
    # both exchange modes walk the same trajectory (split-K atomics leave ~1e-5 noise in the gradients)
    d = float((res[True][0]["flat"] - res[False][0]["flat"]).abs().max())
    assert d < 5e-5, "overlapped vs plain all-reduce: parameters differ by %.3e" % d
    g0, g1 = res[True][0]["summed"], res[False][0]["summed"]
    assert float((g0 - g1).abs().max()) <= 1e-3 * float(g1.abs().max())
