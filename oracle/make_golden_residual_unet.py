"""TEST INFRASTRUCTURE -- generates tests/golden/residual_unet_small.{npz,json} by running the UNMODIFIED reference
`ResidualUNet` (nnunet/network_architecture/generic_modular_residual_UNet.py:273-318; residual decoder :142-270) in the
build container (needs /root/reference; stub recipe in oracle/ref_import.py).

    python oracle/make_golden_residual_unet.py

Reference ResidualUNet (base 8, encoder blocks (1,2,2), decoder blocks (2,1), first stage kernel (1,3,3) / pool (1,1,1),
47 heads with bias, deep supervision) on a (2,1,8,16,16) synthetic CT batch: state_dict (norm parameters perturbed so that
nothing hides behind a zero / unit affine), input, DS targets, logits of the 2 outputs, the reference MultiTalent loss
triple and every parameter gradient.
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
CFG = dict(base=8, blocks_enc=(1, 2, 2), blocks_dec=(2, 1),
           pool=[[1, 1, 1], [1, 2, 2], [2, 2, 2]], convk=[[1, 3, 3], [3, 3, 3], [3, 3, 3]],
           patch=(8, 16, 16), tasks=("Task017_AbdominalOrganSegmentation", "Task064_KiTS_labelsFixed"))


def main():
    ref_import.install()
    ref_import.init_gloo_single()
    torch.set_num_threads(1)
    from nnunet.network_architecture.generic_modular_residual_UNet import ResidualUNet, get_default_network_config
    from nnunet.network_architecture.initialization import InitWeights_He
    torch.manual_seed(0)
    cfg = get_default_network_config(3, None, norm_type="in")
    net = ResidualUNet(1, CFG["base"], CFG["blocks_enc"], 2, CFG["pool"], CFG["convk"], cfg, 47, CFG["blocks_dec"], True,
                       False, 320, InitWeights_He(1e-2))
    g = torch.Generator().manual_seed(13)
    with torch.no_grad():
        for n, p in net.named_parameters():
            is_norm = "norm" in n or n.endswith(("downsample_skip.1.weight", "downsample_skip.1.bias"))
            if is_norm and n.endswith(".weight"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif (is_norm and n.endswith(".bias")) or (("_output" in n) and n.endswith(".bias")):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    rng = np.random.RandomState(987)
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
    w = np.array([2 / 3, 1 / 3])
    out = net(torch.from_numpy(x))
    assert len(out) == len(targets) == 2, (len(out), len(targets))
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
    np.savez_compressed(os.path.join(GOLD, "residual_unet_small.npz"), **blob)
    with open(os.path.join(GOLD, "residual_unet_small.json"), "w") as f:
        json.dump({k: (list(v) if isinstance(v, tuple) else v) for k, v in CFG.items()} |
                  {"valid_regions": [list(v) for v in valid], "torch": torch.__version__,
                   "state_dict_keys": list(net.state_dict().keys())}, f, indent=1)
    print("wrote", os.path.getsize(os.path.join(GOLD, "residual_unet_small.npz")))


if __name__ == "__main__":
    main()
