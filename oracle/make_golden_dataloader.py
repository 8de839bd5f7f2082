"""TEST INFRASTRUCTURE -- generates tests/golden/dataloader_small.npz by running the UNMODIFIED reference
`DataLoader3D.generate_train_batch` (nnunet/training/dataloading/dataset_loading.py:154-380) in the build container
(needs /root/reference; stub recipe in oracle/ref_import.py, incl. the restated `SlimDataLoaderBase` constructor of the
un-vendored batchgenerators).

    python oracle/make_golden_dataloader.py

Four tiny synthetic cases (one smaller than the patch in every axis, one with no foreground) are written as .npy files
the way `unpack_dataset` leaves them; `np.random.seed(7)`; three batches of 3 with pad_mode 'constant', pad_sides
(2, 0, 4), 1/sqrt(n) sampling probabilities and 34 % foreground oversampling are drawn and stored next to the cases.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CFG = dict(patch_size=(14, 20, 20), final_patch_size=(10, 16, 16), batch_size=3, oversample=0.34, pad_sides=(2, 0, 4),
           seed=7, n_batches=3)


def make_cases():
    rng = np.random.RandomState(3)
    shapes = {"003_a": (18, 30, 26), "003_b": (9, 12, 14), "017_c": (22, 24, 40), "006_d": (16, 21, 22)}
    cases = {}
    for k, sh in shapes.items():
        data = rng.randn(1, *sh).astype(np.float32)
        lab = np.zeros((1,) + sh, dtype=np.float32)
        if k != "006_d":          # one case without any foreground
            for l in (1, 2):
                c = [rng.randint(2, s - 2) for s in sh]
                lab[0, c[0] - 2:c[0] + 2, c[1] - 2:c[1] + 3, c[2] - 1:c[2] + 2] = l
        lab[0, 0, :2] = -1        # nnU-Net's "outside the nonzero mask" marker
        cl = {l: np.argwhere(lab[0] == l) for l in (1, 2)}
        cases[k] = (np.concatenate([data, lab]), {'class_locations': cl})
    return cases


def main():
    ref_import.install()
    from nnunet.training.dataloading.dataset_loading import DataLoader3D
    cases = make_cases()
    tmp = tempfile.mkdtemp(prefix="mtb200_dl_")
    dataset = {}
    for k, (arr, props) in cases.items():
        np.save(os.path.join(tmp, k + ".npy"), arr)
        dataset[k] = {'data_file': os.path.join(tmp, k + ".npz"), 'properties': props}
    keys = list(dataset.keys())
    ids = [k.split('_')[0] for k in keys]
    probs = np.array([1 / (ids.count(i) ** 0.5) for i in ids])    # MultiTalent_Trainer_DDP.py:629-633
    probs = probs / probs.sum()
    dl = DataLoader3D(dataset, CFG["patch_size"], CFG["final_patch_size"], CFG["batch_size"], False,
                      oversample_foreground_percent=CFG["oversample"], pad_mode="constant", pad_sides=CFG["pad_sides"],
                      memmap_mode='r', sampling_probabilities=probs)
    np.random.seed(CFG["seed"])
    blob = {"probs": probs}
    for k, (arr, props) in cases.items():
        blob["case/" + k] = arr
        for l, v in props['class_locations'].items():
            blob["loc/%s/%d" % (k, l)] = v
    for i in range(CFG["n_batches"]):
        b = dl.generate_train_batch()
        blob["batch%d/data" % i] = b['data']
        blob["batch%d/seg" % i] = b['seg']
        blob["batch%d/keys" % i] = np.array([str(k) for k in b['keys']])
    np.savez_compressed(os.path.join(GOLD, "dataloader_small.npz"), **blob)
    print("wrote", os.path.getsize(os.path.join(GOLD, "dataloader_small.npz")), [list(blob["batch%d/keys" % i]) for i in range(3)])


if __name__ == "__main__":
    main()
