"""Debug helper: repeat ONE training step from identical parameters and report, per parameter tensor, how much the
gradient differs from the first repetition (relative to the tensor's largest gradient entry)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multitalent_b200.engine import bump_weights_epoch
from multitalent_b200.plans import default_plans
from multitalent_b200.synthetic import synthetic_batch
from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp

dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[sys.argv[1] if len(sys.argv) > 1 else "fp32"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
patch = (16, 32, 32)
plans = default_plans(patch_size=patch, batch_size=2)
plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
plans['plans_per_stage'][1]['conv_kernel_sizes'] = [[3, 3, 3]] * 4
tr = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False, native_dtype=dtype)
torch.manual_seed(0)
tr.initialize(True)
batch = synthetic_batch(patch, 2, 0, tr.deep_supervision_scales)
data = torch.from_numpy(batch['data']).cuda()
tgt = [torch.from_numpy(t).cuda() for t in batch['target']]
valid = [p['valid_regions'] for p in batch['properties']]
flat0 = tr.arena.flat.clone()
names = [n for n, _ in tr.network.named_parameters()]
# oracle gradients on the CPU (the reference algorithm restated, oracle/unet_oracle.py)
from oracle import unet_oracle as O
sd = {k: v.detach().cpu().clone() for k, v in tr.network.state_dict().items()}
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
out = O.generic_unet_forward(torch.from_numpy(batch['data']), sdg, tr.net_num_pool_op_kernel_sizes, tr.net_conv_kernel_sizes)
lo = O.multitalent_loss(out, [torch.from_numpy(t) for t in batch['target']], valid, tr.ds_loss_weights)
lo[0].backward()
ograd = [sdg[n].grad for n in names]
ref = None
for rep in range(reps):
    tr.arena.flat.copy_(flat0)
    tr.arena.mom.zero_()
    tr.arena.first = True
    bump_weights_epoch()
    tr.lr = 0.0  # keep the parameters: only the gradients matter here
    l, _, _ = tr.train_step(data, tgt, valid, True)
    torch.cuda.synchronize()
    grads = [p.grad.detach().clone() for p in tr.network.parameters()]
    relo = [float((g.cpu() - r).abs().max() / (r.abs().max() + 1e-30)) if float(r.abs().max()) > 1e-6 else 0.0
            for g, r in zip(grads, ograd)]
    wo = np.argsort(relo)[::-1][:2]
    print("rep %d vs ORACLE: " % rep + "; ".join("%s %.2e" % (names[i], relo[i]) for i in wo))
    if ref is None:
        ref = grads
        continue
    rel = [float((g - r).abs().max() / (r.abs().max() + 1e-30)) for g, r in zip(grads, ref)]
    order = np.argsort(rel)[::-1][:3]
    print("rep %d loss %.7f: " % (rep, float(l)) + "; ".join("%s %.2e" % (names[i], rel[i]) for i in order))
