#!/usr/bin/env python
"""Headline benchmark: 3D patches/s of the MultiTalent training step (Generic_UNet 3d_fullres, 192x160x128, bs 4 per
GPU, 13-dataset multi-head loss, clip + Nesterov SGD) on N B200s -- BASELINE.json `configs[1]`.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run by the driver)
    python bench.py --net resenc ...                         (BASELINE.json configs[2]: FabiansUNet, resenc plan, bs 4)
    python bench.py --dtype fp16 ...                         (configs[1] as worded: fp16 + dynamic loss scaling)
    python bench.py --impl reference ...                     (the reference algorithm's CPU arm: oracle port on host cores)

Prints ONE JSON line (rank 0).  `value` = whole-job patches/s with inputs resident in HBM; `e2e` = the same through
`MultiTalent_trainer_ddp.run_iteration` with pinned HOST batches (H2D of data+targets and D2H of the loss inside the timed
region); `roofline` = the dominant CUDA kernel against MEASURED_PEAKS.json plus whole-step fractions, the library
(cuDNN under autocast) arm on the same GPU and the sliding-window inference leg as scalar keys; `cpu_baseline` = the
oracle on host cores.
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
# fwd + dgrad + wgrad GFLOP per 192x160x128 patch, true channel counts
TRAIN_GFLOP_PER_PATCH = {"generic": 4769.7, "resenc": 7973.6}
FWD_GFLOP_PER_PATCH = {"generic": 1592.0, "resenc": 2660.0}
WORKLOAD = {
    "generic": "Generic_UNet 3d_fullres (MultiTalent_bs4 plan) training step: fwd + 13-dataset multi-head BCE+Dice loss "
               "+ bwd + clip12 + Nesterov SGD",
    "resenc": "FabiansUNet residual-encoder U-Net (MultiTalent_resenc_bs4 plan) training step: fwd + 13-dataset "
              "multi-head BCE+Dice loss + bwd + clip12 + Nesterov SGD",
}
GENERIC_POOL = [[2, 2, 2]] * 4 + [[1, 2, 2]]
GENERIC_CONVK = [[3, 3, 3]] * 6


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


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on host cores (oracle port only -- nothing of the product is imported here)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_oracle_step_factory(patch, net="generic", seed=0):
    """One training step of the reference algorithm on host cores: oracle forward + MultiTalent loss + backward +
    clip/SGD, bs 1, fp32 (the reference's own CPU path, restated)."""
    from oracle import unet_oracle as O
    if net == "generic":
        pool, convk = GENERIC_POOL, GENERIC_CONVK
        sd0 = O.generic_unet_random_state_dict(pool, convk, seed=seed)
        scales = [[1, 1, 1]] + [list(s) for s in 1 / np.cumprod(np.vstack(pool), axis=0)][:-1]

        def fwd(x, sd):
            return O.generic_unet_forward(x, sd, pool, convk)
        n_out = 5
    else:
        pool = [[1, 1, 1], [1, 2, 2]] + [[2, 2, 2]] * 4
        convk = [[1, 3, 3]] + [[3, 3, 3]] * 5
        be, bd = (1, 2, 3, 4, 4, 4), (1, 1, 1, 1, 1)
        sd0 = O.fabians_unet_random_state_dict(be, pool, convk, bd, seed=seed)
        scales = [[1, 1, 1]] + [list(s) for s in 1 / np.cumprod(np.vstack(pool[1:]), axis=0)][:-1]

        def fwd(x, sd):
            return O.fabians_unet_forward(x, sd, be, pool, convk, bd)
        n_out = 5
    names = list(sd0.keys())
    state = {"p": [sd0[n].clone() for n in names], "buf": [None] * len(names)}
    rng = np.random.RandomState(1234)
    task = O.TASK_IDS[6]
    vol, lab = O.synthetic_ct_and_labels(patch, task, rng)
    x = torch.from_numpy(vol[None, None])
    tg = [torch.from_numpy(t) for t in O.downsample_targets(lab[None, None], scales)]
    valid = [O.VALID_REGIONS[task]]
    w = O.multitalent_ds_loss_weights(n_out)

    def step():
        sd = {n: p.clone().requires_grad_(True) for n, p in zip(names, state["p"])}
        out = fwd(x, sd)
        l, _, _ = O.multitalent_loss(out, tg, valid, w)
        l.backward()
        state["p"], state["buf"], _ = O.clip_and_sgd_step(state["p"], [sd[n].grad for n in names], state["buf"], 1e-2)
        return float(l.detach())
    return step


def cpu_sample_patch(net):
    # resenc has 6 resolution levels (down to 1/16, 1/32, 1/32): the sample must stay > 1 voxel at the bottom
    return (96, 96, 64) if net == "generic" else (64, 128, 96)


def run_reference(args, rank, out=sys.stdout):
    """`--impl reference`: the reference's algorithm on the box's host cores (oracle port; the reference is pure Python
    over torch CPU ops, so there is no separate oracle/_ref binary)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample of the workload per step; if K+W steps of that would not finish within ~4 minutes on this host
    # (timed on the first step), fall back to a smaller sample
    sample_patch = cpu_sample_patch(args.net)
    step = cpu_oracle_step_factory(sample_patch, args.net)
    t_probe = time.perf_counter()
    step()
    t_probe = time.perf_counter() - t_probe
    done_warm = 1
    if t_probe * (args.steps + args.warmup) > 240.0:
        sample_patch = (48, 64, 64) if args.net == "generic" else (32, 64, 64)
        step = cpu_oracle_step_factory(sample_patch, args.net)
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
            "config": {"workload": WORKLOAD[args.net] + " -- CPU sample", "net": args.net,
                       "patch": list(sample_patch), "batch_per_step": 1},
            "cpu_baseline": {"value": value, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# library arm on the same GPU: the reference's ops (cuDNN / ATen under autocast) -- "the real bar", SURVEY.md 2a / 8d
# ----------------------------------------------------------------------------------------------------------------------
def gpu_reference_leg(sd, trainer, data, target, valid, amp_dtype, steps, warmup, net):
    """patches/s of the reference's training step through the library on this GPU: oracle port on CUDA tensors =
    F.conv3d / F.instance_norm / F.leaky_relu (cuDNN, cudnn.benchmark) under torch.autocast(amp_dtype), the reference's
    Python loss loop, GradScaler, clip 12, torch.optim.SGD(nesterov).  Same batch, same warm-up discipline, CUDA events.
    Timed twice: NCDHW tensors (what the reference does) and channels_last_3d (the library at its best)."""
    from oracle.gpu_reference import GpuReference
    sp = trainer.plans['plans_per_stage'][trainer.stage]
    res = {}
    for tag, cl in (("", False), ("_channels_last", True)):
        try:
            if net == "generic":
                ref = GpuReference(sd, trainer.net_num_pool_op_kernel_sizes, trainer.net_conv_kernel_sizes, amp_dtype,
                                   channels_last=cl, ds_loss_weights=trainer.ds_loss_weights)
            else:
                ref = GpuReference(sd, sp['pool_op_kernel_sizes'], sp['conv_kernel_sizes'], amp_dtype, channels_last=cl,
                                   arch="fabians", blocks_enc=sp['num_blocks_encoder'],
                                   blocks_dec=sp['num_blocks_decoder'], ds_loss_weights=trainer.ds_loss_weights)
            for _ in range(warmup):
                l = ref.train_step(data, target, valid)[0]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                l = ref.train_step(data, target, valid)[0]
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res["cudnn%s_ms_per_step" % tag] = ms
            res["cudnn%s_patches_s" % tag] = data.shape[0] / (ms * 1e-3)
            res["cudnn%s_loss" % tag] = float(l)
            res["cudnn%s_peak_mem_gb" % tag] = torch.cuda.max_memory_allocated() / 2 ** 30
        except Exception as exc:  # an out-of-memory library arm must not take the native numbers down with it
            res["cudnn%s_error" % tag] = "%s: %s" % (type(exc).__name__, str(exc)[:160])
        ref = None
        torch.cuda.empty_cache()
    res["cudnn_steps"], res["cudnn_warmup"] = steps, warmup
    res["cudnn_mode"] = "torch %s, cuDNN %s, cudnn.benchmark, autocast(%s), GradScaler, python loss loop" % (
        torch.__version__, torch.backends.cudnn.version(), str(amp_dtype).replace("torch.", ""))
    return res


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
    ap.add_argument("--settle-steps", type=int, default=30,
                    help="untimed steps before the W warm-up steps (clock / memory-pool settling on a fresh box)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--net", default="generic", choices=["generic", "resenc"],
                    help="generic = BASELINE.json configs[1]; resenc = configs[2] (FabiansUNet, resenc plan, bs 4)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--patch", type=int, nargs=3, default=list(FULL_PATCH))
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (ncu passes only)")
    ap.add_argument("--no-infer", action="store_true", help="skip the sliding-window inference leg")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the library (cuDNN autocast) arm")
    ap.add_argument("--infer-vol", type=int, nargs=3, default=[512, 512, 512],
                    help="synthetic volume of the inference leg (BASELINE.json configs[3]: 512^3 = 210 tiles at step 0.5)")
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
    from multitalent_b200.training.network_training.MultiTalent_meets_resenc import MultiTalent_trainer_resenc_ddp

    assert torch.cuda.is_available(), "bench.py (native arm) needs a GPU; there is no CPU fallback"
    L.lib()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    patch = tuple(args.patch)
    cls = MultiTalent_trainer_ddp if args.net == "generic" else MultiTalent_trainer_resenc_ddp
    tr = cls(default_plans(args.net, patch_size=patch, batch_size=args.batch), 0, local_rank,
             native_dtype=dt, init_distributed=world > 1)
    torch.manual_seed(0)  # same initial weights on every rank (they are broadcast anyway)
    tr.initialize(True)
    tr.network._engine.impl = args.kernel_impl
    dev = torch.device("cuda", local_rank)

    batch = synthetic_batch(patch, args.batch, rank, tr.deep_supervision_scales)
    valid = [p['valid_regions'] for p in batch['properties']]
    host_data = torch.from_numpy(batch['data']).pin_memory()
    host_tgt = [torch.from_numpy(t).pin_memory() for t in batch['target']]
    d_data = host_data.to(dev)
    d_tgt = [t.to(dev) for t in host_tgt]
    h2d = host_data.numel() * 4 + sum(t.numel() * 4 for t in host_tgt)
    sd0 = {k: v.detach().clone() for k, v in tr.network.state_dict().items()}  # for the library arm: same start

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)  # NVML attached before the warm-up; samples only while `active` is set
    sampler.start()
    # settle phase (reported as config.settle_steps; same count on every rank): a fresh box occasionally ran the first
    # seconds of its first process 10 - 25 % slow (memory pools, tensor-map and module first use, power state) -- 36.2 /
    # 38.1 / 41.5 ms against 32.6 on the same box seconds later.  These steps are untimed, like the W warm-up steps that
    # follow them.
    for _ in range(max(0, args.settle_steps)):
        tr.train_step(d_data, d_tgt, valid, True)
    barrier()
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
    ksum, ms_profiled, shapes, kernels = {}, None, {}, {}
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
        kernels = kp.per_cuda_kernel()
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

    # ---- library arm (rank 0 of a 1-GPU run): the reference's ops under autocast on the same batch, same GPU
    cudnn = {}
    if world == 1 and not args.no_gpu_reference and args.dtype != "fp32":
        tr._prefetched = {}
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        cudnn = gpu_reference_leg(sd0, tr, d_data, d_tgt, valid, dt, min(args.steps, 8), min(max(args.warmup, 2), 3),
                                  args.net)

    # ---- inference leg (BASELINE.json metric: "train+infer", configs[3]): sliding-window predict_3D of one synthetic
    #      512^3 volume per rank (replicas only, no collective), patch 192x160x128, step 0.5, Gaussian weighting, no
    #      mirroring.  `infer_patches_s`: host volume in, probabilities + segmentation left on the device;
    #      `infer_e2e_patches_s`: the reference contract (numpy in -> numpy out, i.e. including the D2H of the
    #      [47, X, Y, Z] fp32 probabilities).
    infer = {}
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
        pps = world * ntiles / (float(inf_ms.item()) * 1e-3)
        infer = {"infer_patches_s": pps, "infer_seconds_per_volume": float(inf_ms.item()) * 1e-3,
                 "infer_volume": "x".join(str(v) for v in ivol), "infer_tiles": ntiles,
                 "infer_tile_batch": int(getattr(net, "inference_tile_batch", 8)), "infer_mirroring": False,
                 "infer_gpu_launches": int(inf_launches),
                 "infer_tflops": pps / world * FWD_GFLOP_PER_PATCH[args.net] * 1e-3,
                 "infer_e2e_patches_s": world * ntiles / float(inf_e2e_s.item()),
                 "infer_e2e_seconds_per_volume": float(inf_e2e_s.item()),
                 "infer_h2d_bytes": int(vol.nbytes), "infer_d2h_bytes": d2h}

    if rank != 0:
        return
    pk = peaks()
    patches_per_step = args.batch * world
    vox_frac = float(np.prod(patch)) / float(np.prod(FULL_PATCH))
    value = patches_per_step / (ms_per_step * 1e-3)
    e2e_value = patches_per_step / (e2e_ms_per_step * 1e-3) if e2e_ms_per_step else None

    # roofline of the dominant kernel = the CUDA kernel (as dispatched inside the library, mtb200_last_kernel) with the
    # largest share of the device time inside the timed region: algorithmic FLOPs (true channel counts) of its launches
    # / their measured duration = average per launch over average per launch.
    total_kernel_ms = sum(v["ms"] for v in kernels.values())
    step_tflops = TRAIN_GFLOP_PER_PATCH[args.net] * vox_frac * args.batch / ms_per_step  # GFLOP/ms == TFLOP/s per GPU
    roof = {}
    traffic_tab = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu --set full
    if os.path.exists(tp):
        with open(tp) as f:
            traffic_tab = json.load(f)
    if kernels:
        top = max(kernels, key=lambda k: kernels[k]["ms"])
        t = kernels[top]
        conv_ms = sum(v["ms"] for k, v in ksum.items() if k.startswith("conv_"))
        conv_fl = sum(v["flops"] for k, v in ksum.items() if k.startswith("conv_"))
        if t["flops"] > 0:
            achieved = t["flops"] / (t["ms"] * 1e-3) / 1e12
            peak, unit, bound = pk["bf16_tflops_sustained"], "TFLOP/s", "tensor"
        else:  # a streaming kernel on top: report it against HBM instead (bytes are not tracked per launch -> None)
            achieved, peak, unit, bound = None, pk["hbm_gbs"], "GB/s", "hbm"
        tr_entry = traffic_tab.get("kernel:" + top)
        roof = {"kernel": top + "_kernel", "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak if achieved else None,
                "traffic": tr_entry.get("dram_bytes_per_launch") if isinstance(tr_entry, dict) else tr_entry,
                "algorithmic_bytes_per_launch": tr_entry.get("algorithmic_bytes_per_launch")
                if isinstance(tr_entry, dict) else None,
                "peak_source": pk["source"] + " (sustained bf16: kernels timed inside a long step)",
                "launches_per_step": t["launches"] / args.steps,
                "share_of_kernel_time": t["ms"] / total_kernel_ms,
                "flops_per_launch": t["flops"] / t["launches"], "ms_per_launch_avg": t["ms"] / t["launches"],
                "ms_per_step": t["ms"] / args.steps,
                # the whole path, not its best slice
                "conv_family_tflops": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms else None,
                "family_frac": conv_fl / (conv_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"] if conv_ms else None,
                "conv_share_of_kernel_time": conv_ms / total_kernel_ms if total_kernel_ms else None}
    roof.update({"step_tflops": step_tflops, "step_frac_sustained": step_tflops / pk["bf16_tflops_sustained"],
                 "step_frac_burst": step_tflops / pk["bf16_tflops"],
                 "target_ms_per_step_at_half_burst": TRAIN_GFLOP_PER_PATCH[args.net] * vox_frac * args.batch /
                 (0.5 * pk["bf16_tflops"])})
    for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])[:12]:  # per-CUDA-kernel ms per step (scalars)
        roof["ms_%s" % k] = round(v["ms"] / args.steps, 4)
    # the HBM-bound passes against the measured copy bandwidth: algorithmic bytes (2 reads + 1 write, 1 + 1, 2 reads of the
    # 16-bit tensors) / their measured durations
    for k in ("in_bwd_apply_tma", "norm_act_tma", "in_bwd_reduce_tma"):
        v = kernels.get(k)
        if v and v.get("bytes") and v["ms"] > 0:
            roof["gbs_%s" % k] = round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)
            roof["hbm_frac_%s" % k] = round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], 3)
    if cudnn:
        roof.update(cudnn)
        best = max([cudnn.get("cudnn_patches_s") or 0.0, cudnn.get("cudnn_channels_last_patches_s") or 0.0])
        if best > 0:
            roof["vs_cudnn"] = value / best
            roof["vs_cudnn_ncdhw"] = value / cudnn["cudnn_patches_s"] if cudnn.get("cudnn_patches_s") else None
    if infer:
        roof.update({k: infer[k] for k in ("infer_patches_s", "infer_seconds_per_volume", "infer_volume", "infer_tiles",
                                          "infer_tflops", "infer_gpu_launches")})
        roof["infer_frac_burst"] = infer["infer_tflops"] / pk["bf16_tflops"]
    line = {"metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD[args.net], "net": args.net,
                       "patch": "x".join(str(v) for v in patch), "batch_per_gpu": args.batch,
                       "global_batch": patches_per_step,
                       "parallelism": "dp%d" % world, "l2": "working set (activations > 8 GB) far exceeds the 126 MB L2",
                       "loss": loss_val, "loss_scaling": tr.loss_scaling_description(),
                       "settle_steps": int(max(0, args.settle_steps))},
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 12,
                    "ms_per_step": e2e_ms_per_step},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "ms_per_step_with_per_launch_events": ms_profiled,
            "per_launch_events_note": "second pass, weight-gradient stream serialised into the main stream",
            "ms_each_step": [round(m, 3) for m in ms_each],
            "top_shapes": [{"kernel": k[0], "shape": k[1], "launches": v["launches"],
                            "ms_per_step": round(v["ms"] / args.steps, 4),
                            "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None}
                           for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:80]],
            "cuda_kernels": {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                                 "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] and v["ms"] else None}
                             for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
            "kernels": {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                            "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] and v["ms"] else None}
                        for k, v in sorted(ksum.items(), key=lambda kv: -kv[1]["ms"])}}
    if infer:
        line["e2e"].update({k: infer[k] for k in ("infer_e2e_patches_s", "infer_e2e_seconds_per_volume",
                                                  "infer_h2d_bytes", "infer_d2h_bytes")})
        line["config"].update({"infer_volume": infer["infer_volume"], "infer_tiles": infer["infer_tiles"],
                               "infer_tile_batch": infer["infer_tile_batch"], "infer_mirroring": False})
        line["infer"] = infer
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample_patch = cpu_sample_patch(args.net)
        frac = float(np.prod(sample_patch)) / float(np.prod(FULL_PATCH))
        step = cpu_oracle_step_factory(sample_patch, args.net)
        step()
        c0 = time.perf_counter()
        nrep = 3
        for _ in range(nrep):
            step()
        cdt = (time.perf_counter() - c0) / nrep
        line["cpu_baseline"] = {"value": frac / cdt, "unit": "patches/s", "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": "oracle fwd+loss+bwd+SGD, bs1, patch %dx%dx%d (%.3f of 192x160x128), %d timed "
                                          "steps after 1 warm-up, scaled by voxel count" % (sample_patch + (frac, nrep))}
    print(json.dumps(line), file=out, flush=True)


if __name__ == "__main__":
    main()
