#!/usr/bin/env python
"""Headline benchmark: 3D patches/s of the MultiTalent training step (Generic_UNet 3d_fullres, 192x160x128, bs 4 per
GPU, 13-dataset multi-head loss, clip + Nesterov SGD) on N B200s -- BASELINE.json `configs[1]`.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run by the driver)
    python bench.py --impl reference ...                     (the reference algorithm's CPU arm: oracle port on host cores)

Prints ONE JSON line (rank 0).  `value` = whole-job patches/s with inputs resident in HBM; `e2e` = the same through
`MultiTalent_trainer_ddp.run_iteration` with pinned HOST batches (H2D of data+targets and D2H of the loss inside the timed
region); `roofline` = the dominant kernel family against MEASURED_PEAKS.json; `cpu_baseline` = the oracle on host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "3D patches/sec (192x160x128, bs4) training step"
FULL_PATCH = (192, 160, 128)
TRAIN_GFLOP_PER_PATCH = 4769.7   # BASELINE.md section 2 (fwd + dgrad + wgrad, true channel counts)


def cuda_kernel_name(tag, shape_repr):
    """CUDA kernel that `mtb200_conv_taps` / `mtb200_wgrad_taps` dispatch this (family, shape) to in the 16-bit
    tensor-core mode -- restates the rules of csrc/conv_umma.cu::conv_taps_umma / wgrad_taps_umma for the report (the ncu
    launch list under profiles/ is the ground truth).  None if the shape is not a convolution."""
    try:
        cin, cout, grid, taps, istr, ostr = eval(shape_repr, {"__builtins__": {}})
        unit = tuple(istr) == (1, 1, 1) and tuple(ostr) == (1, 1, 1)
        w = grid[2]
        if cin == 1:
            return "conv_c1_wgrad_kernel" if tag == "conv_wgrad" else "conv_c1_fwd_kernel"
        if tag == "conv_wgrad":
            if unit and w >= 48 and (taps >= 9 or taps == 1) and (cin == 16 or cin % 32 == 0):
                return "wgrad_line_umma_kernel"
            return "wgrad_taps_umma_kernel"
        if taps == 1 and unit:
            return "conv_pw_umma_kernel"
        if tuple(ostr) != (1, 1, 1) and tuple(istr) == (1, 1, 1) and grid[2] >= 64 and cin <= 64 and cout <= 64:
            return "conv_gm_umma_kernel"  # output lattice problems at the top level: strided dgrad, ConvTranspose fwd
        if unit and taps >= 9 and w >= 72 and cin <= 64 and cout % 32 == 0:
            return "conv_line_umma_kernel"
        return "conv_taps_umma_kernel"
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md "clocks" line).  In-process NVML
    (initialised before the warm-up): forking `nvidia-smi` next to the launch loop costs a process spawn plus an NVML
    attach per sample, which stalled the launching thread for tens of ms inside a 0.25 s timed region.  `nvidia-smi`
    remains the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self._stop_evt, self.active = [], threading.Event(), threading.Event()
        self.nv = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _sample(self):
        if self.nv is not None:
            nv = self.nv
            mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                    nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
            return mhz, self.max_mhz, [n for n, b in zip(self.NAMES, bits) if r & b]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [s.strip() for s in out.strip().split(",")]
        return float(parts[0]), float(parts[1]), [n for n, v in zip(self.NAMES, parts[2:6])
                                                  if v.lower().startswith("active")]

    def run(self):
        while not self._stop_evt.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(self._sample())
                except Exception:
                    pass
            self._stop_evt.wait(self.period if self.nv is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [s[0] for s in self.samples]
        mx = [s[1] for s in self.samples]
        reasons = sorted({n for s in self.samples for n in s[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples),
                "source": "nvml (in-process)" if self.nv is not None else "nvidia-smi"}


def cpu_oracle_step_factory(patch, seed=0):
    """One training step of the reference algorithm on host cores: oracle forward + MultiTalent loss + backward +
    clip/SGD, bs 1, fp32 (the reference's own CPU path, restated)."""
    from oracle import unet_oracle as O
    from multitalent_b200.network_architecture.generic_UNet import Generic_UNet, InitWeights_He
    from torch import nn
    pool = [[2, 2, 2]] * 4 + [[1, 2, 2]]
    convk = [[3, 3, 3]] * 6
    torch.manual_seed(seed)
    net = Generic_UNet(1, 30, 47, 5, 2, 2, nn.Conv3d, nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d,
                       {'p': 0, 'inplace': True}, nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True}, True, False,
                       lambda x: x, InitWeights_He(1e-2), pool, convk, False, True, True)
    names = [n for n, _ in net.named_parameters()]
    state = {"p": [p.detach().clone() for _, p in net.named_parameters()], "buf": [None] * len(names)}
    rng = np.random.RandomState(1234)
    task = O.TASK_IDS[6]
    vol, lab = O.synthetic_ct_and_labels(patch, task, rng)
    x = torch.from_numpy(vol[None, None])
    scales = [[1, 1, 1], [.5] * 3, [.25] * 3, [.125] * 3, [1 / 16] * 3]
    tg = [torch.from_numpy(t) for t in O.downsample_targets(lab[None, None], scales)]
    valid = [O.VALID_REGIONS[task]]
    w = O.multitalent_ds_loss_weights(5)

    def step():
        sd = {n: p.clone().requires_grad_(True) for n, p in zip(names, state["p"])}
        out = O.generic_unet_forward(x, sd, pool, convk)
        l, _, _ = O.multitalent_loss(out, tg, valid, w)
        l.backward()
        state["p"], state["buf"], _ = O.clip_and_sgd_step(state["p"], [sd[n].grad for n in names], state["buf"], 1e-2)
        return float(l)
    return step


def run_reference(args, rank, out=sys.stdout):
    """`--impl reference`: the reference's algorithm on the box's host cores (oracle port; the reference is pure Python
    over torch CPU ops, so there is no separate oracle/_ref binary)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample of the workload: 0.15 of the benchmark patch per step; if K+W steps of that would not finish within
    # ~4 minutes on this host (timed on the first step), fall back to a 48x64x64 sample (0.05 of the patch)
    sample_patch = (96, 96, 64)
    step = cpu_oracle_step_factory(sample_patch)
    t_probe = time.perf_counter()
    step()
    t_probe = time.perf_counter() - t_probe
    done_warm = 1
    if t_probe * (args.steps + args.warmup) > 240.0:
        sample_patch = (48, 64, 64)
        step = cpu_oracle_step_factory(sample_patch)
        done_warm = 0
    frac = float(np.prod(sample_patch)) / float(np.prod(FULL_PATCH))
    for _ in range(max(0, args.warmup - done_warm)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = frac / dt
    sample = ("oracle port (torch CPU fp32): fwd + MultiTalent loss + bwd + clip/SGD, bs1, patch %dx%dx%d = %.3f of a "
              "192x160x128 patch per step, scaled by voxel count" % (sample_patch + (frac,)))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Generic_UNet 3d_fullres training step, 13-dataset multi-head loss, CPU sample",
                       "patch": list(sample_patch), "batch_per_step": 1},
            "cpu_baseline": {"value": value, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """Libraries print to stdout (NCCL's version banner under torchrun): the contract is ONE JSON line there.  Point fd 1
    at stderr for the whole run and return a file object on the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--patch", type=int, nargs=3, default=list(FULL_PATCH))
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (ncu passes only)")
    ap.add_argument("--no-infer", action="store_true", help="skip the sliding-window inference leg")
    ap.add_argument("--infer-vol", type=int, nargs=3, default=[384, 320, 256],
                    help="synthetic volume of the inference leg (27 tiles of 192x160x128 at step 0.5)")
    ap.add_argument("--no-profile", action="store_true", help="no per-launch CUDA events in the timed region")
    ap.add_argument("--kernel-impl", type=int, default=0, help="0 auto, 1 force CUDA-core kernels, 2 force tcgen05")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, out)
        return

    import torch.distributed as dist
    from multitalent_b200 import _lib as L
    from multitalent_b200.plans import default_plans
    from multitalent_b200.synthetic import synthetic_batch
    from multitalent_b200.training.network_training.MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp

    assert torch.cuda.is_available(), "bench.py (native arm) needs a GPU; there is no CPU fallback"
    L.lib()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    patch = tuple(args.patch)
    tr = MultiTalent_trainer_ddp(default_plans(patch_size=patch, batch_size=args.batch), 0, local_rank,
                                 native_dtype=dt, init_distributed=world > 1)
    torch.manual_seed(0)  # same initial weights on every rank (they are broadcast anyway)
    tr.initialize(True)
    tr.network._engine.impl = args.kernel_impl
    if args.dtype == "fp16":
        tr.loss_scale = 4096.0  # static loss scale (the reference uses a dynamic GradScaler, MT:350-354)
    dev = torch.device("cuda", local_rank)

    batch = synthetic_batch(patch, args.batch, rank, tr.deep_supervision_scales)
    valid = [p['valid_regions'] for p in batch['properties']]
    host_data = torch.from_numpy(batch['data']).pin_memory()
    host_tgt = [torch.from_numpy(t).pin_memory() for t in batch['target']]
    d_data = host_data.to(dev)
    d_tgt = [t.to(dev) for t in host_tgt]
    h2d = host_data.numel() * 4 + sum(t.numel() * 4 for t in host_tgt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)  # NVML attached before the warm-up; samples only while `active` is set
    sampler.start()
    for _ in range(args.warmup):
        tr.train_step(d_data, d_tgt, valid, True)
    barrier()

    # ---- timed region: K steps, inputs resident in HBM (no per-launch instrumentation: this is `value`)
    n0 = L.launch_count
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    sampler.active.set()
    marks[0].record()
    for i in range(args.steps):
        l, ce, dc = tr.train_step(d_data, d_tgt, valid, True)
        marks[i + 1].record()
    barrier()
    sampler.active.clear()
    e0, e1 = marks[0], marks[-1]
    ms_each = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    launches = L.launch_count - n0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    loss_val = float(l.item())

    # ---- the same K steps again with one CUDA-event pair around every launch (on the launching stream): per-kernel
    #      durations for the roofline section.  The events cost ~3 % of a step, which is why `value` is taken above.
    ksum, ms_profiled, shapes = {}, None, {}
    if not args.no_profile:
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # per-kernel durations are taken with the weight-gradient stream folded back into the main stream: two kernels
        # sharing the SMs would each be charged the other's time
        eng = tr.network._engine
        overlap, eng.overlap_wgrad = eng.overlap_wgrad, False
        with L.KernelProfile() as kp:
            barrier()
            p0.record()
            for _ in range(args.steps):
                tr.train_step(d_data, d_tgt, valid, True)
            p1.record()
            barrier()
        eng.overlap_wgrad = overlap
        ksum = kp.summary()
        ms_profiled = p0.elapsed_time(p1) / args.steps
        for tag, info, kms, fl, _nb in kp.per_launch():  # group launches by (family, problem shape)
            g = shapes.setdefault((tag, repr(info)), {"launches": 0, "ms": 0.0, "flops": 0.0})
            g["launches"] += 1
            g["ms"] += kms
            g["flops"] += fl
    clocks = sampler.stop()

    # ---- end to end through run_iteration: pinned host batch -> H2D -> step -> D2H of (l, ce, dc)
    def gen():
        while True:
            yield {'data': host_data, 'target': host_tgt, 'properties': batch['properties']}
    g = gen()
    e2e_ms_per_step = None
    if not args.no_e2e:
        tr.run_iteration(g, True)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            tr.run_iteration(g, True)
        t1.record()
        barrier()
        e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e_ms_per_step = float(e2e_ms.item()) / args.steps

    # ---- inference leg (BASELINE.json metric: "train+infer"): sliding-window predict_3D of one synthetic volume per
    #      rank (replicas only, no collective), patch 192x160x128, step 0.5, Gaussian weighting, no mirroring.
    #      `value`: host volume in, probabilities + segmentation left on the device; `e2e`: the reference contract
    #      (numpy in -> numpy out, i.e. including the D2H of the [47, X, Y, Z] fp32 probabilities).
    infer = None
    if not args.no_infer and patch == FULL_PATCH:
        del d_data, d_tgt
        tr._prefetched = {}
        torch.cuda.empty_cache()
        ivol = tuple(args.infer_vol)
        vol = np.random.RandomState(7 + rank).randn(1, *ivol).astype(np.float32)
        net = tr.network
        steps_xyz = net._compute_steps_for_sliding_window(patch, ivol, 0.5)
        ntiles = len(steps_xyz[0]) * len(steps_xyz[1]) * len(steps_xyz[2])
        kw = dict(do_mirroring=False, mirror_axes=(0, 1, 2), use_sliding_window=True, step_size=0.5, use_gaussian=True,
                  verbose=False)
        ds, mode = net.do_ds, net.training
        net.do_ds = False
        net.eval()
        try:
            ikw = dict(kw, patch_size=patch, regions_class_order=tuple(range(47)), return_device_tensors=True)
            seg, prob = net.predict_3D(vol, **ikw)  # warm-up: the same volume once (allocator, Gaussian map, tile batch)
            del seg, prob
            barrier()
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_inf = L.launch_count
            i0.record()
            seg, prob = net.predict_3D(vol, **ikw)
            i1.record()
            barrier()
            inf_launches = L.launch_count - n_inf
            inf_ms = torch.tensor([i0.elapsed_time(i1)], dtype=torch.float64, device=dev)
            del seg, prob
            torch.cuda.empty_cache()
            # steady state of a multi-volume job: the page-locked result buffers of the previous volume are recycled
            seg_np, prob_np = tr.predict_preprocessed_data_return_seg_and_softmax(vol, do_mirroring=False, verbose=False)
            del seg_np, prob_np
            torch.cuda.synchronize()
            t_host = time.perf_counter()
            seg_np, prob_np = tr.predict_preprocessed_data_return_seg_and_softmax(vol, do_mirroring=False, verbose=False)
            torch.cuda.synchronize()
            inf_e2e_s = torch.tensor([time.perf_counter() - t_host], dtype=torch.float64, device=dev)
            d2h = int(seg_np.nbytes + prob_np.nbytes)
            del seg_np, prob_np
        finally:
            net.train(mode)
            net.do_ds = ds
        if world > 1:
            dist.all_reduce(inf_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(inf_e2e_s, op=dist.ReduceOp.MAX)
        infer = {"metric": "3D patches/sec (192x160x128) sliding-window inference", "unit": "patches/s",
                 "value": world * ntiles / (float(inf_ms.item()) * 1e-3),
                 "seconds_per_volume": float(inf_ms.item()) * 1e-3, "volume": list(ivol), "tiles_per_volume": ntiles,
                 "volumes": world, "tile_batch": int(getattr(net, "inference_tile_batch", 4)), "mirroring": False,
                 "gpu_launches": int(inf_launches),
                 "includes": "H2D of the volume, per tile: gather + forward + sigmoid*Gaussian scatter-add, normalise + "
                             "threshold; results left in HBM",
                 "e2e": {"value": world * ntiles / float(inf_e2e_s.item()), "unit": "patches/s",
                         "seconds_per_volume": float(inf_e2e_s.item()), "h2d_bytes": int(vol.nbytes),
                         "d2h_bytes": d2h, "timing": "host wall clock around predict_preprocessed_data_return_seg_and_"
                                                       "softmax (numpy in, numpy out), second volume of a run"}}

    if rank != 0:
        return
    pk = peaks()
    patches_per_step = args.batch * world
    vox_frac = float(np.prod(patch)) / float(np.prod(FULL_PATCH))
    value = patches_per_step / (ms_per_step * 1e-3)
    e2e_value = patches_per_step / (e2e_ms_per_step * 1e-3) if e2e_ms_per_step else None

    # roofline of the dominant kernel = the (kernel family, problem shape) group with the most device time inside the
    # timed region; algorithmic FLOPs per launch (true channel counts) / average measured launch duration
    total_kernel_ms = sum(v["ms"] for v in ksum.values())
    conv_shapes = {k: v for k, v in shapes.items() if k[0].startswith("conv_") and v["flops"] > 0}
    roof = None
    traffic_tab = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu --set full
    if os.path.exists(tp):
        with open(tp) as f:
            traffic_tab = json.load(f)
    if conv_shapes:
        top = max(conv_shapes, key=lambda k: conv_shapes[k]["ms"])
        t = conv_shapes[top]
        achieved = t["flops"] / (t["ms"] * 1e-3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        fam = ksum[top[0]]
        roof = {"kernel": top[0], "cuda_kernel": cuda_kernel_name(top[0], top[1]) if args.dtype != "fp32" else None,
                "shape(Cin_p,Cout_p,grid,taps,in_stride,out_stride)": top[1], "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic_tab.get("%s %s" % top), "peak_source": pk["source"] + " (sustained bf16)",
                "launches": t["launches"], "share_of_kernel_time": t["ms"] / total_kernel_ms,
                "flops_per_launch": t["flops"] / t["launches"], "ms_per_launch_avg": t["ms"] / t["launches"],
                "family": {"launches": fam["launches"], "share_of_kernel_time": fam["ms"] / total_kernel_ms,
                           "achieved": fam["flops"] / (fam["ms"] * 1e-3) / 1e12}}
    step_tflops = TRAIN_GFLOP_PER_PATCH * vox_frac * args.batch / ms_per_step  # GFLOP/ms == TFLOP/s
    line = {"metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "Generic_UNet 3d_fullres (MultiTalent_bs4 plan) training step: fwd + 13-dataset "
                                   "multi-head BCE+Dice loss + bwd + clip12 + Nesterov SGD",
                       "patch": list(patch), "batch_per_gpu": args.batch, "global_batch": patches_per_step,
                       "parallelism": "dp%d" % world, "l2": "working set (activations > 8 GB) far exceeds the 126 MB L2",
                       "loss": loss_val},
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 12,
                    "ms_per_step": e2e_ms_per_step},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "infer": infer,
            "ms_per_step_with_per_launch_events": ms_profiled,
            "per_launch_events_note": "second pass, weight-gradient stream serialised into the main stream",
            "ms_each_step": [round(m, 3) for m in ms_each],
            "top_shapes": [{"kernel": k[0], "shape": k[1], "launches": v["launches"],
                            "ms_per_step": round(v["ms"] / args.steps, 4),
                            "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None}
                           for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:40]],
            "step_tflops_algorithmic": step_tflops / world * world if world == 1 else step_tflops,
            "step_frac_of_bf16_peak": step_tflops / pk["bf16_tflops_sustained"],
            "kernels": {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                            "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] and v["ms"] else None}
                        for k, v in sorted(ksum.items(), key=lambda kv: -kv[1]["ms"])}}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample_patch = (96, 96, 64)
        frac = float(np.prod(sample_patch)) / float(np.prod(FULL_PATCH))
        step = cpu_oracle_step_factory(sample_patch)
        step()
        c0 = time.perf_counter()
        nrep = 2
        for _ in range(nrep):
            step()
        cdt = (time.perf_counter() - c0) / nrep
        line["cpu_baseline"] = {"value": frac / cdt, "unit": "patches/s", "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": "oracle fwd+loss+bwd+SGD, bs1, patch 96x96x64 (0.15 of 192x160x128), %d timed "
                                          "steps after 1 warm-up, scaled by voxel count" % nrep}
    print(json.dumps(line), file=out, flush=True)


if __name__ == "__main__":
    main()
