"""Debug helper: repeat the 3-step run of tests/test_gpu_parity.py::test_run_iteration_prefetch_matches_unprefetched and
report, per step, which parameters differ between repetitions."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multitalent_b200.plans import default_plans
from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp

blob = np.load(os.path.join(ROOT, "tests/golden/generic_small.npz"))
meta = json.load(open(os.path.join(ROOT, "tests/golden/generic_small.json")))
plans = default_plans(patch_size=(8, 16, 16), batch_size=2)
plans['plans_per_stage'][1]['pool_op_kernel_sizes'] = meta["pool"]
plans['plans_per_stage'][1]['conv_kernel_sizes'] = meta["convk"]
plans['base_num_features'] = meta["base"]
sd = {k[len("param/"):]: torch.from_numpy(blob[k]) for k in blob.files if k.startswith("param/")}
rng = np.random.RandomState(5)
batches = []
for i in range(3):
    x = (blob["x"] + 0.1 * i * rng.randn(*blob["x"].shape)).astype(np.float32)
    batches.append({'data': torch.from_numpy(x) if os.environ.get("UNPINNED") else torch.from_numpy(x).pin_memory(),
                    'target': [(lambda t: t if os.environ.get("UNPINNED") else t.pin_memory())(torch.from_numpy(np.roll(blob["target_%d" % k], i, axis=-1).copy())) for k in range(3)],
                    'properties': [{'valid_regions': tuple(v)} for v in meta["valid_regions"]]})
runs = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    prefetch = bool(rep % 2)
    tr = MultiTalent_trainer_ddp(plans, 0, 0, init_distributed=False)
    tr.initialize(True)
    tr.prefetch_batches = prefetch
    tr.load_checkpoint_ram({'state_dict': dict(sd), 'epoch': 0})
    g = iter(batches)
    snaps, losses = [], []
    names = [n for n, _ in tr.network.named_parameters()]
    for s in range(3):
        losses.append([float(v) for v in tr.run_iteration(g, True)])
        snaps.append([p.detach().cpu().clone() for p in tr.network.parameters()])
    runs.append((prefetch, losses, snaps))
    if rep:
        for s in range(3):
            d = [float((a - b).abs().max()) for a, b in zip(snaps[s], runs[0][2][s])]
            w = int(np.argmax(d))
            if max(d) > 1e-6:
                dd = (snaps[s][w] - runs[0][2][s][w]).abs()
                nz = (dd > 1e-6).nonzero()
                print("   shape", tuple(dd.shape), "elements > 1e-6:", nz.shape[0], "first:", nz[:12].tolist(),
                      "values", [round(float(dd[tuple(i)]), 7) for i in nz[:6]])
            print("rep %d prefetch=%d step %d: max param diff vs rep0 %.3e at %s ; loss diff %.3e" %
                  (rep, prefetch, s, max(d), names[w], max(abs(a - b) for a, b in zip(losses[s], runs[0][1][s]))))
