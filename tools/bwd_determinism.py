"""Debug helper: run the network forward + backward twice in one process with identical inputs / seed gradients and
report the first intermediate buffers (forward features, gradient buffers, parameter gradients) that differ."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multitalent_b200.engine import Tape
from multitalent_b200.plans import default_plans
from multitalent_b200.synthetic import synthetic_batch
from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp

if os.environ.get("NAN_EMPTY"):  # poison every fresh CUDA allocation: reads of never-written memory surface as NaN
    _e, _el = torch.empty, torch.empty_like
    def _empty(*a, **k):
        t = _e(*a, **k)
        if t.is_cuda and t.is_floating_point():
            t.fill_(float("nan"))
        return t
    def _empty_like(*a, **k):
        t = _el(*a, **k)
        if t.is_cuda and t.is_floating_point():
            t.fill_(float("nan"))
        return t
    torch.empty, torch.empty_like = _empty, _empty_like

dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[sys.argv[1] if len(sys.argv) > 1 else "fp32"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
patch = (16, 32, 32)
plans = default_plans(patch_size=patch, batch_size=2)
plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
tr = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False, native_dtype=dtype)
torch.manual_seed(0)
tr.initialize(True)
net, eng = tr.network, tr.network._engine
if os.environ.get("NO_WS"):
    eng.overlap_wgrad = False
batch = synthetic_batch(patch, 2, 0, tr.deep_supervision_scales)
data = torch.from_numpy(batch['data']).cuda()
names = [n for n, _ in net.named_parameters()]
seeds = None
runs = []
for rep in range(reps):
    tr.arena.zero_grad()
    tape = Tape()
    eng.begin_step()
    feats = net._native_forward(data, tape)
    if seeds is None:
        g = torch.Generator(device="cuda").manual_seed(1)
        seeds = [torch.randn(f.as_ncdhw().shape, device="cuda", generator=g) for f in feats]
    for f, s in zip(feats, seeds):
        eng.seed_grad(tape, f, s)
    eng.run_backward(tape)
    torch.cuda.synchronize()
    rec = {"out": [f.buf.float().clone() for f in feats],
           "gbuf": [v.float().clone() for v in tape.grad_bufs.values()],
           "pgrad": [p.grad.detach().clone() for p in net.parameters()]}
    runs.append(rec)
    if os.environ.get("NAN_EMPTY"):
        nan_o = [i for i, t in enumerate(rec["out"]) if torch.isnan(t[..., :47]).any()]
        nan_g = [(i, tuple(t.shape), int(torch.isnan(t).sum())) for i, t in enumerate(rec["gbuf"]) if torch.isnan(t).any()]
        nan_p = [names[i] for i, t in enumerate(rec["pgrad"]) if torch.isnan(t).any()]
        print("rep %d NaN check: outputs %s | grad buffers %s | param grads %s" % (rep, nan_o, nan_g[:8], nan_p[:8]))
    if rep:
        r0 = runs[0]
        def rel(a, b):
            return float((a - b).abs().max() / (b.abs().max() + 1e-30))
        do = [rel(a, b) for a, b in zip(rec["out"], r0["out"])]
        dg = [rel(a, b) for a, b in zip(rec["gbuf"], r0["gbuf"])]
        dp = [rel(a, b) for a, b in zip(rec["pgrad"], r0["pgrad"])]
        bad_g = [(i, tuple(rec["gbuf"][i].shape), "%.1e" % d) for i, d in enumerate(dg) if d > 1e-5]
        bad_p = [(names[i], "%.1e" % d) for i, d in enumerate(dp) if d > 1e-5]
        print("rep %d: max out diff %.1e | grad buffers (creation order) differing: %s | param grads differing: %s" %
              (rep, max(do), bad_g[:6], bad_p[:6]))
