"""TEST INFRASTRUCTURE -- generates tests/golden/resenc_small.npz and dc_ce_small.npz by running the UNMODIFIED reference
in the build container (needs /root/reference; stub recipe in oracle/ref_import.py).

    python oracle/make_golden_resenc.py

  resenc_small.npz   reference FabiansUNet (base 8, encoder blocks (1,2,2,2), first stage kernel (1,3,3) / pool (1,1,1),
                     47 heads with bias) on a (2,1,8,16,16) synthetic CT batch: state_dict (norm parameters perturbed so
                     that the zero-initialised norm2 does not hide conv2), input, DS targets, logits of the 3 outputs,
                     the reference MultiTalent loss triple and every parameter gradient
  dc_ce_small.npz    reference DC_and_CE_loss (batch_dice, smooth 1e-5, do_bg False) under MultipleOutputLoss2 on random
                     3-scale logits: loss and d(loss)/d(logits)
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CFG = dict(base=8, blocks_enc=(1, 2, 2, 2), blocks_dec=(1, 1, 1),
           pool=[[1, 1, 1], [1, 2, 2], [2, 2, 2], [2, 2, 2]], convk=[[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]],
           patch=(8, 16, 16), tasks=("Task017_AbdominalOrganSegmentation", "Task003_Liver"))


def main():
    ref_import.install()
    ref_import.init_gloo_single()
    torch.set_num_threads(1)
    from nnunet.network_architecture.generic_modular_residual_UNet import FabiansUNet, get_default_network_config
    from nnunet.network_architecture.initialization import InitWeights_He
    from nnunet.training.loss_functions.dice_loss import DC_and_CE_loss
    from nnunet.training.loss_functions.deep_supervision import MultipleOutputLoss2

    # ---- residual-encoder U-Net
    torch.manual_seed(0)
    cfg = get_default_network_config(3, None, norm_type="in")
    net = FabiansUNet(1, CFG["base"], CFG["blocks_enc"], 2, CFG["pool"], CFG["convk"], cfg, 47, CFG["blocks_dec"], True,
                      False, 320, InitWeights_He(1e-2))
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "norm" in n and n.endswith(".weight") or n.endswith("downsample_skip.1.weight"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif ("norm" in n and n.endswith(".bias")) or n.endswith("downsample_skip.1.bias") or \
                    ("deep_supervision_outputs" in n and n.endswith(".bias")):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    rng = np.random.RandomState(4321)
    vols, labs = [], []
    for t in CFG["tasks"]:
        v, l = O.synthetic_ct_and_labels(CFG["patch"], t, rng)
        vols.append(v)
        labs.append(l)
    x = np.stack(vols)[:, None].astype(np.float32)
    lab = np.stack(labs)[:, None].astype(np.float32)
    scales = [[1, 1, 1]] + [list(s) for s in 1 / np.cumprod(np.vstack(CFG["pool"][1:]), axis=0)][:-1]
    targets = O.downsample_targets(lab, scales)
    valid = [O.VALID_REGIONS[t] for t in CFG["tasks"]]
    n = len(CFG["pool"])
    w = np.array([1 / (2 ** i) for i in range(n)])
    w[n - 1] = 0
    w = w / w.sum()
    out = net(torch.from_numpy(x))
    assert len(out) == len(targets) == 3
    l, ce, dc = ref_import.reference_compute_loss(out, [torch.from_numpy(t) for t in targets], valid, w)
    net.zero_grad()
    l.backward()
    blob = {"x": x, "ds_loss_weights": w.astype(np.float64),
            "loss": np.array([l.item(), ce.item(), dc.item()], dtype=np.float64)}
    for i, t in enumerate(targets):
        blob["target_%d" % i] = t
    for i, o in enumerate(out):
        blob["logits_%d" % i] = o.detach().numpy()
    for nme, p in net.named_parameters():
        blob["param/" + nme] = p.detach().numpy()
        blob["grad/" + nme] = p.grad.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "resenc_small.npz"), **blob)
    with open(os.path.join(GOLD, "resenc_small.json"), "w") as f:
        json.dump({k: (list(v) if isinstance(v, tuple) else v) for k, v in CFG.items()} |
                  {"valid_regions": [list(v) for v in valid], "torch": torch.__version__}, f, indent=1)

    # ---- generic softmax Dice + CE with deep supervision
    torch.manual_seed(5)
    C = 5
    shapes = [(2, C, 8, 12, 16), (2, C, 4, 6, 8), (2, C, 2, 3, 4)]
    zs = [torch.randn(s, requires_grad=True) for s in shapes]
    ts = [torch.randint(0, C, (s[0], 1) + s[2:]).float() for s in shapes]
    wds = np.array([4 / 7, 2 / 7, 1 / 7])
    loss_fn = MultipleOutputLoss2(DC_and_CE_loss({'batch_dice': True, 'smooth': 1e-5, 'do_bg': False}, {}), wds)
    lv = loss_fn(zs, ts)
    lv.backward()
    blob = {"loss": np.array([lv.item()]), "weights": wds}
    for i in range(3):
        blob["logits_%d" % i] = zs[i].detach().numpy()
        blob["target_%d" % i] = ts[i].numpy()
        blob["grad_%d" % i] = zs[i].grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "dc_ce_small.npz"), **blob)
    print("wrote", {f: os.path.getsize(os.path.join(GOLD, f)) for f in ("resenc_small.npz", "dc_ce_small.npz")})


if __name__ == "__main__":
    main()
