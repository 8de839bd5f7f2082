"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python oracle/make_golden.py          (needs /root/reference; see oracle/ref_import.py for the stub recipe)

Fixtures (small, committed):
  tables.json          the Task100_MultiTalent tables dumped from the reference module
  generic_small.npz    a 3-pool Generic_UNet (base 8 features, 47 heads) on a (2,1,8,16,16) synthetic CT batch:
                       reference state_dict, input, DS targets, valid regions, reference logits (3 scales), the
                       reference MultiTalent loss triple (gloo world 1) and all parameter gradients of d(loss)
  sliding_small.npz    reference predict_3D (tiled, Gaussian, 8-way mirroring, sigmoid) on a (1,12,24,28) volume with the
                       same network: segmentation + sub-sampled probabilities
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

SMALL = dict(pool=[[2, 2, 2], [2, 2, 2], [1, 2, 2]], convk=[[3, 3, 3]] * 4, base=8, patch=(8, 16, 16),
             tasks=("Task017_AbdominalOrganSegmentation", "Task003_Liver"))


def small_inputs():
    rng = np.random.RandomState(1234)
    vols, labs = [], []
    for t in SMALL["tasks"]:
        v, l = O.synthetic_ct_and_labels(SMALL["patch"], t, rng)
        vols.append(v)
        labs.append(l)
    x = np.stack(vols)[:, None].astype(np.float32)
    lab = np.stack(labs)[:, None].astype(np.float32)
    scales = [[1, 1, 1]] + [list(s) for s in 1 / np.cumprod(np.vstack(SMALL["pool"]), axis=0)][:-1]
    targets = O.downsample_targets(lab, scales)
    return x, targets


def main():
    ref_import.install()
    ref_import.init_gloo_single()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order

    # ---- tables
    from nnunet.dataset_conversion import Task100_MultiTalent as T
    tables = {
        "task_ids": T.MultiTalent_task_ids,
        "task_label_maps": {k: [list(v[0]), list(v[1])] for k, v in T.MultiTalent_task_label_maps.items()},
        "labels": {str(k): v for k, v in T.MultiTalent_labels.items()},
        "regions": [[k, list(v)] for k, v in T.MultiTalent_regions.items()],
        "regions_class_order": {k: list(v) for k, v in T.MultiTalent_regions_class_order.items()},
        "region_output_idx_mapping": T.MultiTalent_region_output_idx_mapping,
        "valid_regions": {k: list(v) for k, v in T.MultiTalent_valid_regions.items()},
    }
    with open(os.path.join(GOLD, "tables.json"), "w") as f:
        json.dump(tables, f, indent=1, sort_keys=True)

    # ---- small Generic_UNet: forward, loss, gradients
    net = ref_import.build_reference_generic_unet(1, SMALL["base"], 47, SMALL["pool"], SMALL["convk"], seed=0)
    # make the affine norm parameters and biases non-trivial so the test is not blind to them
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("instnorm.weight"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif n.endswith("instnorm.bias") or n.endswith("conv.bias"):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    x, targets = small_inputs()
    xt = torch.from_numpy(x)
    tg = [torch.from_numpy(t) for t in targets]
    valid = [O.VALID_REGIONS[t] for t in SMALL["tasks"]]
    w = O.multitalent_ds_loss_weights(len(SMALL["pool"]))
    out = net(xt)
    l, ce, dc = ref_import.reference_compute_loss(out, tg, valid, w)
    net.zero_grad()
    l.backward()
    blob = {"x": x, "ds_loss_weights": w.astype(np.float64),
            "loss": np.array([l.item(), ce.item(), dc.item()], dtype=np.float64)}
    for i, t in enumerate(targets):
        blob["target_%d" % i] = t
    for i, o in enumerate(out):
        blob["logits_%d" % i] = o.detach().numpy()
    for n, p in net.named_parameters():
        blob["param/" + n] = p.detach().numpy()
        blob["grad/" + n] = p.grad.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "generic_small.npz"), **blob)
    with open(os.path.join(GOLD, "generic_small.json"), "w") as f:
        json.dump({"pool": SMALL["pool"], "convk": SMALL["convk"], "base": SMALL["base"], "tasks": SMALL["tasks"],
                   "valid_regions": [list(v) for v in valid], "torch": torch.__version__}, f, indent=1)

    # ---- sliding window
    rng = np.random.RandomState(99)
    vol, _ = O.synthetic_ct_and_labels((12, 24, 28), SMALL["tasks"][0], rng)
    net.eval()
    net.do_ds = False
    with torch.no_grad():
        seg, prob = net.predict_3D(vol[None], do_mirroring=True, mirror_axes=(0, 1, 2), use_sliding_window=True,
                                   step_size=0.5, patch_size=SMALL["patch"], regions_class_order=tuple(range(47)),
                                   use_gaussian=True, all_in_gpu=False, verbose=False, mixed_precision=False)
        seg2, prob2 = net.predict_3D(vol[None], do_mirroring=False, use_sliding_window=True, step_size=0.5,
                                     patch_size=SMALL["patch"], regions_class_order=tuple(range(47)),
                                     use_gaussian=True, all_in_gpu=False, verbose=False, mixed_precision=False)
    np.savez_compressed(os.path.join(GOLD, "sliding_small.npz"), vol=vol, seg_mirror=seg.astype(np.float32),
                        prob_mirror_sub=prob[:, ::2, ::2, ::2].astype(np.float32),
                        seg_nomirror=seg2.astype(np.float32), prob_nomirror_sub=prob2[:, ::2, ::2, ::2].astype(np.float32))
    print("wrote fixtures to", GOLD, {f: os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD)})


if __name__ == "__main__":
    main()
