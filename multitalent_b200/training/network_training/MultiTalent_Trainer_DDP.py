"""`MultiTalent_trainer_ddp` (alias `nnUNetTrainerV2_MultiTalent`) -- the hot-path half of
nnunet/training/network_training/custom_trainers/MultiTalent/MultiTalent/MultiTalent_Trainer_DDP.py:30-127, 324-370,
544-623 and of the nnUNetTrainerV2_DDP pieces it inherits (nnUNetTrainerV2_DDP.py:50-133, 601-634;
nnUNetTrainerV2.py:131-170, 393-408), on the native kernels.

Same constructor signature, same hook names (`initialize`, `initialize_network`, `compute_loss`, `run_iteration`,
`predict_preprocessed_data_return_seg_and_softmax`, `maybe_update_lr`), same return contracts.  Everything around the
hot path that the reference trainer also does (data loaders, augmentation, validation export, plotting, checkpoint
files) is out of scope and stays in the reference; INTEGRATION.md shows how the reference
trainer picks up the native network + loss with a 3-line subclass.

Two step modes:
  * autograd mode (`run_iteration`): exactly the reference's statement sequence -- `network(data)`, `compute_loss`,
    `backward()`, clip, SGD -- where the network and the loss are single autograd nodes backed by the kernels.  Works
    under torch DDP unchanged.
  * `flat_optimizer=True` (default): parameters and gradients live in one fp32 arena each; the clip-norm and the
    Nesterov-SGD update are two kernels over the arena and the DDP gradient exchange is one NCCL all-reduce of it.
"""
import os
import pickle
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

from ... import _lib as L
from ...dataset_conversion.Task100_MultiTalent import MultiTalent_regions
from ...engine import bump_weights_epoch, pad_channels
from ...network_architecture.generic_UNet import Generic_UNet, InitWeights_He
from ...plans import default_plans
from ..loss_functions.multitalent_loss import head_windows, multitalent_loss
from ..online_evaluation import OnlineEvaluationMixin
from ..validation import ValidationMixin


def poly_lr(epoch, max_epochs, initial_lr, exponent=0.9):
    """nnunet/training/learning_rate/poly_lr.py:16-17."""
    return initial_lr * (1 - epoch / max_epochs) ** exponent


class FlatArena:
    """All parameters (and their gradients) of a module as views into one contiguous fp32 buffer each."""

    def __init__(self, module: nn.Module):
        params = [p for p in module.parameters()]

        # 1-D parameters (bias, norm weight / bias) get a slot padded to the kernels' channel padding: the engine
        # views them in place (engine._padded) and accumulates their gradients in place (engine.direct_grad)
        slot = self._slot
        n = sum(slot(p) for p in params)
        dev = params[0].device
        self.params = params
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            p.grad = self.grad[off:off + k].view_as(p.data)
            p._mtb_direct_grad = True
            if p.dim() == 1:
                p._mtb_padded_len = slot(p)
            off += slot(p)
        self.n = n
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.first = True

    def zero_grad(self):
        self.grad.zero_()

    def step(self, lr, momentum, weight_decay, max_norm, inv_scale=1.0, scaler=None):
        """clip_grad_norm_(max_norm) + SGD(nesterov) over the arena.  Gradients are multiplied by `inv_scale` (1 / world
        size of the summed all-reduce) and, with a `DeviceGradScaler`, by 1 / its current loss scale (read on the
        device); a non-finite gradient norm skips the update (GradScaler.step semantics) and the scaler then backs off."""
        st = L.stream_ptr()
        self.sumsq.zero_()
        L.call("mtb200_sumsq", L.ptr(self.grad), self.n, L.ptr(self.sumsq), st)
        L.call("mtb200_sgd_step", L.ptr(self.flat), L.ptr(self.grad), L.ptr(self.mom), self.n, L.ptr(self.sumsq),
               float(inv_scale), float(max_norm), float(lr), float(momentum), float(weight_decay), int(self.first),
               L.ptr(scaler.state) if scaler is not None else None, st)
        if scaler is not None:
            scaler.update(self.sumsq)
        self.first = False
        # the kernel updated the arena behind torch's version counters: invalidate the packed-weight caches
        bump_weights_epoch()

    # ---- torch.optim.SGD-format state (checkpoints interchangeable with the reference, network_trainer.py:256-286) ---
    def sgd_state_dict(self, lr, momentum, weight_decay):
        opt = torch.optim.SGD(self.params, lr, weight_decay=weight_decay, momentum=momentum, nesterov=True)
        sd = opt.state_dict()
        if not self.first:
            off = 0
            for i, p in enumerate(self.params):
                k = p.numel()
                sd['state'][i] = {'momentum_buffer': self.mom[off:off + k].view_as(p.data).detach().cpu().clone()}
                off += self._slot(p)
        return sd

    def load_sgd_state_dict(self, sd):
        state = sd.get('state', {})
        self.mom.zero_()
        off, any_buf = 0, False
        for i, p in enumerate(self.params):
            k = p.numel()
            st = state.get(i, state.get(str(i)))
            buf = st.get('momentum_buffer') if st else None
            if buf is not None:
                self.mom[off:off + k].copy_(buf.reshape(-1).to(self.mom.device, torch.float32))
                any_buf = True
            off += self._slot(p)
        self.first = not any_buf

    @staticmethod
    def _slot(p):
        return pad_channels(p.numel()) if p.dim() == 1 else (p.numel() + 3) // 4 * 4


class DeviceGradScaler:
    """torch.cuda.amp.GradScaler semantics (MultiTalent_Trainer_DDP.py:349-354: scale -> backward -> unscale_ -> clip ->
    step -> update) with the state on the device, so the arena step needs no host synchronisation: `state` = fp32
    {scale, growth_tracker, found_inf of the last step, skipped steps}.  `scale(l)` multiplies by the device scalar; the
    SGD kernel unscales by it and skips the update when the gradient norm is not finite; `update` = GradScaler.update
    (backoff x0.5 on inf/nan, growth x2 after `growth_interval` clean steps).  `state_dict` uses GradScaler's keys."""

    def __init__(self, device, init_scale=2.0 ** 16, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self.state = torch.tensor([init_scale, 0.0, 0.0, 0.0], dtype=torch.float32, device=device)

    def scale(self, loss):
        return loss * self.state[0]

    def update(self, sumsq):
        L.call("mtb200_loss_scale_update", L.ptr(sumsq), L.ptr(self.state), float(self.growth_factor),
               float(self.backoff_factor), int(self.growth_interval), L.stream_ptr())

    def get_scale(self):
        return float(self.state[0].item())

    def state_dict(self):
        st = self.state.cpu()
        return {"scale": float(st[0]), "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": int(st[1])}

    def load_state_dict(self, sd):
        self.growth_factor, self.backoff_factor = float(sd["growth_factor"]), float(sd["backoff_factor"])
        self.growth_interval = int(sd["growth_interval"])
        self.state[0] = float(sd["scale"])
        self.state[1] = float(sd.get("_growth_tracker", 0))


class MultiTalent_trainer_ddp(OnlineEvaluationMixin, ValidationMixin):
    def __init__(self, plans_file, fold, local_rank, output_folder=None, dataset_directory=None, batch_dice=True,
                 stage=None, unpack_data=True, deterministic=True, distribute_batch_size=False, fp16=False,
                 native_dtype=None, flat_optimizer=True, init_distributed=True):
        """Positional signature of MultiTalent_Trainer_DDP.py:31-32 (the tuple is pickled and replayed by
        model_restore.py:90).  Keyword-only extensions: `native_dtype` (torch.float32 | bfloat16 | float16; default
        fp16 if `fp16` else fp32), `flat_optimizer`, `init_distributed`.  float16 runs as the reference runs it
        (MT:349-354): dynamic loss scaling with GradScaler semantics (`amp_grad_scaler`); bfloat16 / float32 need none."""
        self.init_args = (plans_file, fold, local_rank, output_folder, dataset_directory, batch_dice, stage, unpack_data,
                          deterministic, distribute_batch_size, fp16)
        self.plans_file, self.fold, self.local_rank = plans_file, fold, local_rank
        self.output_folder, self.dataset_directory = output_folder, dataset_directory
        self.batch_dice = True  # forced, MT:33
        self.stage, self.unpack_data, self.deterministic = stage, unpack_data, deterministic
        self.distribute_batch_size, self.fp16 = distribute_batch_size, fp16
        self.native_dtype = native_dtype if native_dtype is not None else (torch.float16 if fp16 else torch.float32)
        self.flat_optimizer = flat_optimizer
        self.regions = MultiTalent_regions
        self.max_num_epochs, self.initial_lr, self.weight_decay = 1000, 1e-2, 3e-5  # nnUNetTrainerV2.py:47-48
        self.num_batches_per_epoch, self.num_val_batches_per_epoch = 250, 50         # network_trainer.py:96-97
        self.epoch = 0
        self.was_initialized = False
        self.plans = None
        self.network = self.optimizer = self.arena = self.amp_grad_scaler = None
        self.dataset_val = None
        self.ds_loss_weights = None
        self.loss_scale = 1.0  # optional STATIC factor on top (experiments); the dynamic scaler is `amp_grad_scaler`
        # network_trainer.py:71-93 bookkeeping that the checkpoint format carries ('plot_stuff', 'best_stuff')
        self.all_tr_losses, self.all_val_losses, self.all_val_losses_tr_mode, self.all_val_eval_metrics = [], [], [], []
        self.best_epoch_based_on_MA_tr_loss = self.best_MA_tr_loss_for_patience = self.best_val_eval_criterion_MA = None
        np.random.seed(local_rank)
        torch.manual_seed(local_rank)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(local_rank)
            torch.cuda.set_device(local_rank)
        # nnUNetTrainerV2_DDP.py:68 -- one process per GPU, env:// rendezvous (launcher provides the env)
        if init_distributed and dist.is_available() and not dist.is_initialized() and "RANK" in os.environ:
            dist.init_process_group(backend='nccl' if torch.cuda.is_available() else 'gloo', init_method='env://')

    # ---- plans -------------------------------------------------------------------------------------------------
    def load_plans_file(self):
        if isinstance(self.plans_file, dict):
            self.plans = self.plans_file
        elif self.plans_file is None:
            self.plans = default_plans()
        else:
            with open(self.plans_file, 'rb') as f:
                self.plans = pickle.load(f)

    def process_plans(self, plans):
        """nnUNetTrainer.py:326-392 (the fields the hot path needs) + MT:48-51 (num_classes := 47 regions)."""
        if self.stage is None:
            self.stage = max(plans['plans_per_stage'].keys())
        sp = plans['plans_per_stage'][self.stage]
        self.plans = plans
        self.batch_size = int(sp['batch_size'])
        self.net_pool_per_axis = sp['num_pool_per_axis']
        self.patch_size = np.array(sp['patch_size']).astype(int)
        self.net_num_pool_op_kernel_sizes = [list(map(int, k)) for k in sp['pool_op_kernel_sizes']]
        self.net_conv_kernel_sizes = [list(map(int, k)) for k in sp['conv_kernel_sizes']]
        self.base_num_features = int(plans['base_num_features'])
        self.num_input_channels = int(plans['num_modalities'])
        self.conv_per_stage = int(plans.get('conv_per_stage', 2))
        self.threeD = len(self.patch_size) == 3
        self.num_classes = len(self.regions)

    def setup_DA_params(self):
        """nnUNetTrainerV2.py:350-351: scales of the deep-supervision targets."""
        self.deep_supervision_scales = [[1, 1, 1]] + list(
            list(i) for i in 1 / np.cumprod(np.vstack(self.net_num_pool_op_kernel_sizes), axis=0))[:-1]

    # ---- initialisation -------------------------------------------------------------------------------------------
    def initialize(self, training=True, force_load_plans=False):
        if self.was_initialized:
            return
        if force_load_plans or self.plans is None:
            self.load_plans_file()
        self.process_plans(self.plans)
        self.setup_DA_params()
        if training:
            n = len(self.net_num_pool_op_kernel_sizes)
            w = np.array([1 / (2 ** i) for i in range(n)])
            w[n - 1] = 0                       # lowest resolution output is not supervised (MT:93-95)
            self.ds_loss_weights = w / w.sum()
        self.initialize_network()
        self.initialize_optimizer_and_scheduler()
        self._wrap_ddp()
        self.was_initialized = True
        self.regions_class_order = list(range(self.num_classes))

    def initialize_network(self):
        """nnUNetTrainerV2.initialize_network (nnUNetTrainerV2.py:131-164) + sigmoid inference nonlinearity (MT:43-46)."""
        assert self.threeD, "MultiTalent is a 3d_fullres configuration"
        self.network = Generic_UNet(self.num_input_channels, self.base_num_features, self.num_classes,
                                    len(self.net_num_pool_op_kernel_sizes), self.conv_per_stage, 2, nn.Conv3d,
                                    nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d,
                                    {'p': 0, 'inplace': True}, nn.LeakyReLU,
                                    {'negative_slope': 1e-2, 'inplace': True}, True, False, lambda x: x,
                                    InitWeights_He(1e-2), self.net_num_pool_op_kernel_sizes,
                                    self.net_conv_kernel_sizes, False, True, True, native_dtype=self.native_dtype)
        if torch.cuda.is_available():
            self.network.cuda()
        self.network.inference_apply_nonlin = nn.Sigmoid()

    def initialize_optimizer_and_scheduler(self):
        """nnUNetTrainerV2.py:166-170: SGD(lr, wd 3e-5, momentum .99, nesterov); no scheduler object (poly LR)."""
        assert self.network is not None, "self.initialize_network must be called first"
        if self.flat_optimizer and next(self.network.parameters()).is_cuda:
            self.arena = FlatArena(self.network)
            self.optimizer = None
        else:
            self.optimizer = torch.optim.SGD(self.network.parameters(), self.initial_lr,
                                             weight_decay=self.weight_decay, momentum=0.99, nesterov=True)
        self.lr = self.initial_lr
        self.lr_scheduler = None
        self._maybe_init_amp()

    def _maybe_init_amp(self):
        """network_trainer.py:400-402: a GradScaler whenever the arithmetic is fp16."""
        if self.amp_grad_scaler is None and self.native_dtype == torch.float16 and torch.cuda.is_available():
            if self.arena is not None:
                self.amp_grad_scaler = DeviceGradScaler(self.arena.flat.device)
            else:
                self.amp_grad_scaler = torch.amp.GradScaler("cuda")

    def loss_scaling_description(self):
        if self.amp_grad_scaler is None:
            return "none" if self.loss_scale == 1.0 else "static x%g" % self.loss_scale
        return "dynamic (GradScaler semantics: init 65536, x0.5 on inf/nan, x2 after 2000 clean steps)"

    def _wrap_ddp(self):
        self.world_size = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if self.world_size > 1:
            with torch.no_grad():  # parameter broadcast at construction, as torch DDP does (MT:121)
                flat = self.arena.flat if self.arena is not None else None
                if flat is not None:
                    dist.broadcast(flat, 0)
                else:
                    for p in self.network.parameters():
                        dist.broadcast(p.data, 0)
        self._setup_overlapped_allreduce()

    # ---- gradient exchange (replaces the DDP reducer of MT:121) --------------------------------------------------------
    def _setup_overlapped_allreduce(self):
        """Arena ranges for the two-part gradient all-reduce.  In the backward pass the gradients of everything but the
        first two encoder stages (> 99 % of the bytes: decoder, bottleneck, deep encoder stages, heads) are final while
        those two widest stages are still being differentiated: that part is all-reduced on a communication stream under
        the rest of the backward pass, the small remainder afterwards.  MTB200_EARLY_ALLREDUCE=0: one all-reduce at the end."""
        self._early_ranges = None
        self._early_issued = False
        net = self.network
        if (self.arena is None or self.world_size <= 1 or os.environ.get("MTB200_EARLY_ALLREDUCE", "1") == "0"
                or not hasattr(net, "conv_blocks_context") or len(net.conv_blocks_context) < 4):
            return
        late = {id(p) for st in list(net.conv_blocks_context)[:2] for p in st.parameters()}
        off, lo, hi = 0, None, None
        for p in self.arena.params:
            k = FlatArena._slot(p)
            if id(p) in late:
                lo = off if lo is None else lo
                hi = off + k
            off += k
        if lo is None:
            return
        self._late_range = (lo, hi)
        self._early_ranges = [(a, b) for a, b in ((0, lo), (hi, self.arena.n)) if b > a]
        self._comm_stream = torch.cuda.Stream()
        net._engine.backward_mark = self._early_allreduce

    def _early_allreduce(self):
        """Backward-pass closure (placed by the network after its second encoder stage): fold the finished weight
        gradients into the arena and all-reduce the early ranges on the communication stream."""
        eng = self.network._engine
        main, comm = torch.cuda.current_stream(), self._comm_stream
        comm.wait_stream(main)
        for st in eng.side_streams_in_use():   # weight-gradient launches of the finished layers
            comm.wait_stream(st)
        with torch.cuda.stream(comm):
            eng._flush_unpack()
            for a, b in self._early_ranges:
                dist.all_reduce(self.arena.grad[a:b])
        self._early_issued = True

    def _finish_allreduce(self):
        if self._early_ranges is not None and self._early_issued:
            self._early_issued = False
            a, b = self._late_range
            dist.all_reduce(self.arena.grad[a:b])
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        else:
            dist.all_reduce(self.arena.grad)

    def maybe_update_lr(self, epoch=None):
        """nnUNetTrainerV2.py:393-408."""
        ep = self.epoch + 1 if epoch is None else epoch
        self.lr = poly_lr(ep, self.max_num_epochs, self.initial_lr, 0.9)
        if self.optimizer is not None:
            self.optimizer.param_groups[0]['lr'] = self.lr

    # ---- the hot path ----------------------------------------------------------------------------------------------
    def compute_loss(self, output, target, valid_regions):
        """MT:544-623 -> (total_loss, total_ce, total_dc).  When the step will be followed by `run_online_evaluation`
        (`_want_hard_stats`), the statistics pass also counts the hard tp / fp / fn of the full-resolution output."""
        hard = {} if getattr(self, "_want_hard_stats", False) else None
        res = multitalent_loss(output, target, valid_regions, self.ds_loss_weights, hard_out=hard,
                               engine=getattr(self.network, "_engine", None))
        self._hard_stats = hard
        return res

    def _stage_batch(self, data_dict):
        """Start the host->device copies of one batch on the copy stream (pinned host memory: asynchronous).  Returns
        (data, target, valid_regions, event after the data copy, event after the target copies): the targets are not
        needed before the loss, so their copy may still be running under the forward pass."""
        data, target = data_dict['data'], data_dict['target']
        valid_regions = [p['valid_regions'] for p in data_dict['properties']]
        data = torch.as_tensor(data)
        target = [torch.as_tensor(t) for t in target]
        ready = tready = None
        if torch.cuda.is_available():
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream()
            host = [t for t in [data] + target if not t.is_cuda]
            if host and all(t.is_pinned() for t in host):
                # no wait on the compute stream: the staging tensors are allocated from the copy stream's own pool
                with torch.cuda.stream(self._copy_stream):
                    data = data.cuda(non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self._copy_stream)
                    target = [t.cuda(non_blocking=True) for t in target]
                    tready = torch.cuda.Event()
                    tready.record(self._copy_stream)
            else:
                data = data.cuda(non_blocking=True)
                target = [t.cuda(non_blocking=True) for t in target]
        return data, target, valid_regions, ready, tready

    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False):
        """MT:324-370.  Returns three numpy scalars (the D2H sync of the reference is kept: it is the contract).

        `prefetch_batches` (default on): the NEXT batch is pulled from `data_generator` and its H2D copy is issued on a
        side stream while this step computes, so the copy of a step's inputs is hidden under the previous step (the
        reference's `to_cuda` copies are synchronous with the step, to_torch.py:18-31).  The generator is therefore
        advanced one batch ahead of the step that consumes it; a batch staged for one generator is only ever handed to
        that generator's next call."""
        pre = getattr(self, "_prefetched", None)
        if pre is None:
            pre = self._prefetched = {}
        entry = pre.pop(id(data_generator), None)
        staged = entry[1] if entry is not None and entry[0] is data_generator else None
        if staged is None:
            staged = self._stage_batch(next(data_generator))
        data, target, valid_regions, ready, tready = staged
        if ready is not None:
            main = torch.cuda.current_stream()
            main.wait_event(ready)
            for t in [data] + target:
                t.record_stream(main)
        l, ce, dc = self.train_step(data, target, valid_regions, do_backprop, target_ready=tready,
                                    keep_output=run_online_evaluation)
        if run_online_evaluation:
            self.run_online_evaluation(self._last_output, target, valid_regions)
            self._last_output = None
        if getattr(self, "prefetch_batches", True) and torch.cuda.is_available():
            try:
                nxt = next(data_generator)
            except StopIteration:
                nxt = None
            if nxt is not None:
                pre[id(data_generator)] = (data_generator, self._stage_batch(nxt))
        res = torch.stack((l.detach(), ce.detach(), dc.detach())).cpu().numpy()
        return res[0], res[1], res[2]

    def train_step(self, data, target, valid_regions, do_backprop=True, target_ready=None, keep_output=False):
        """One optimisation step on device-resident tensors; returns device scalars (no sync).  `target_ready`: CUDA
        event after which `target` may be read (asynchronous H2D copy issued by `run_iteration`)."""
        if self.arena is not None:
            self.arena.zero_grad()
        elif self.optimizer is not None:
            self.optimizer.zero_grad()
        # training steps that do not hand their logits to anybody (no online evaluation): the fusable heads are deferred --
        # computed inside the loss's statistics pass and recomputed inside the fused backward, never written to memory
        eng = getattr(self.network, "_engine", None)
        defer = bool(eng is not None and do_backprop and not keep_output and getattr(eng, "fuse_head", False)
                     and getattr(eng, "defer_head_fwd", False) and not torch.is_tensor(valid_regions)
                     and self.network.training
                     and type(self).compute_loss is MultiTalent_trainer_ddp.compute_loss  # a custom loss needs real logits
                     and head_windows(valid_regions, (int(getattr(self.network, "num_classes", 0)) + 7) // 8 * 8)
                     is not None)
        with torch.set_grad_enabled(do_backprop):
            if eng is not None:
                eng.defer_heads = defer
            try:
                output = self.network(data)
            finally:
                if eng is not None:
                    eng.defer_heads = False
            if target_ready is not None:
                torch.cuda.current_stream().wait_event(target_ready)
            self._want_hard_stats = bool(keep_output)
            l, ce, dc = self.compute_loss(output, target, valid_regions)
            self._want_hard_stats = False
            if keep_output:  # run_online_evaluation reads the highest-resolution logits after the step (MT:367-368)
                self._last_output = tuple(o.detach() for o in output)
            if do_backprop:
                ls = l * self.loss_scale if self.loss_scale != 1.0 else l
                if self.amp_grad_scaler is not None:
                    ls = self.amp_grad_scaler.scale(ls)        # MT:350
                ls.backward()
        if do_backprop:
            if self.arena is not None:
                if self.world_size > 1:
                    self._finish_allreduce()
                inv = 1.0 / (self.loss_scale * self.world_size)
                self.arena.step(self.lr, 0.99, self.weight_decay, 12.0, inv, scaler=self.amp_grad_scaler)
            else:
                params = [p for p in self.network.parameters() if p.grad is not None]
                if self.world_size > 1 or self.loss_scale != 1.0:
                    for p in params:
                        if self.world_size > 1:
                            dist.all_reduce(p.grad)
                        p.grad.div_(self.world_size * self.loss_scale)
                if self.amp_grad_scaler is not None:           # MT:351-354
                    self.amp_grad_scaler.unscale_(self.optimizer)
                    torch.nn.utils.clip_grad_norm_(self.network.parameters(), 12)
                    self.amp_grad_scaler.step(self.optimizer)
                    self.amp_grad_scaler.update()
                else:
                    torch.nn.utils.clip_grad_norm_(self.network.parameters(), 12)
                    self.optimizer.step()
        return l, ce, dc

    def predict_preprocessed_data_return_seg_and_softmax(self, data, do_mirroring=True, mirror_axes=None,
                                                         use_sliding_window=True, step_size=0.5, use_gaussian=True,
                                                         pad_border_mode='constant', pad_kwargs=None, all_in_gpu=False,
                                                         verbose=True, mixed_precision=True, region_vec=None):
        """nnUNetTrainerV2_DDP.py:601-634: do_ds off, eval mode, predict_3D with regions_class_order, restore state."""
        if pad_border_mode == 'constant' and pad_kwargs is None:
            pad_kwargs = {'constant_values': 0}
        if do_mirroring and mirror_axes is None:
            mirror_axes = (0, 1, 2)  # default_data_augmentation.py:70
        net = self.network
        ds, mode = net.do_ds, net.training
        net.do_ds = False
        net.eval()
        try:
            ret = net.predict_3D(data, do_mirroring=do_mirroring, mirror_axes=mirror_axes or (),
                                 use_sliding_window=use_sliding_window, step_size=step_size,
                                 patch_size=tuple(self.patch_size), regions_class_order=self.regions_class_order,
                                 use_gaussian=use_gaussian, pad_border_mode=pad_border_mode, pad_kwargs=pad_kwargs,
                                 all_in_gpu=all_in_gpu, verbose=verbose, mixed_precision=mixed_precision)
        finally:
            net.train(mode)
            net.do_ds = ds
        return ret

    # ---- checkpoints (network_trainer.py:256-286, nnUNetTrainerV2_DDP.py:636-669: key names are the contract) -------
    def save_checkpoint(self, fname, save_optimizer=True):
        """network_trainer.py:256-286: the reference's key set, so that either side can load the other's file.  In arena
        mode the optimizer state is written in torch.optim.SGD's format (momentum buffers per parameter)."""
        sd = {k: v.detach().cpu().clone() for k, v in self.network.state_dict().items()}
        opt_sd = None
        if save_optimizer and self.arena is not None:
            opt_sd = self.arena.sgd_state_dict(self.lr, 0.99, self.weight_decay)
        elif save_optimizer and self.optimizer is not None:
            opt_sd = self.optimizer.state_dict()
        state = {'epoch': self.epoch + 1, 'state_dict': sd, 'optimizer_state_dict': opt_sd,
                 'lr_scheduler_state_dict': None,
                 'plot_stuff': (self.all_tr_losses, self.all_val_losses, self.all_val_losses_tr_mode,
                                self.all_val_eval_metrics),
                 'best_stuff': (self.best_epoch_based_on_MA_tr_loss, self.best_MA_tr_loss_for_patience,
                                self.best_val_eval_criterion_MA)}
        if self.amp_grad_scaler is not None:
            state['amp_grad_scaler'] = self.amp_grad_scaler.state_dict()
        torch.save(state, fname)
        with open(fname + ".pkl", 'wb') as f:
            pickle.dump({'init': self.init_args, 'name': self.__class__.__name__, 'class': str(self.__class__),
                         'plans': self.plans}, f)

    def load_checkpoint(self, fname, train=True):
        """network_trainer.py:337-345."""
        if not self.was_initialized:
            self.initialize(train)
        self.load_checkpoint_ram(torch.load(fname, map_location=torch.device('cpu'), weights_only=False), train)

    def load_checkpoint_ram(self, checkpoint, train=True):
        """nnUNetTrainerV2_DDP.py:636-697: `module.` prefix heuristic (+ the legacy resenc head remap, :656-659),
        GradScaler state, optimizer state (momentum), bookkeeping lists, epoch correction; then the poly learning rate
        of the restored epoch (nnUNetTrainerV2.run_training -> maybe_update_lr(self.epoch))."""
        if not self.was_initialized:
            self.initialize(train)
        cur = self.network.state_dict()
        legacy = {'module.decoder.segmentation_output.weight': 'decoder.deep_supervision_outputs.4.weight',
                  'module.decoder.segmentation_output.bias': 'decoder.deep_supervision_outputs.4.bias'}
        new = {}
        for k, v in checkpoint['state_dict'].items():
            key = k
            if key not in cur:
                key = legacy.get(key, key[7:] if key.startswith('module.') else key)
            new[key] = v
        missing = [k for k in cur if k not in new]
        unexpected = [k for k in new if k not in cur]
        if missing or unexpected:  # nn.Module.load_state_dict(strict=True) semantics
            raise RuntimeError("Error(s) in loading state_dict: missing keys %s, unexpected keys %s" % (missing, unexpected))
        with torch.no_grad():
            for k, v in new.items():
                cur[k].copy_(v)  # in place: parameters may be views of the flat arena
        bump_weights_epoch()
        if self.amp_grad_scaler is not None and train and checkpoint.get('amp_grad_scaler') is not None:
            self.amp_grad_scaler.load_state_dict(checkpoint['amp_grad_scaler'])
        self.epoch = checkpoint.get('epoch', 0)
        if train:
            opt_sd = checkpoint.get('optimizer_state_dict')
            if opt_sd is not None:
                if self.arena is not None:
                    self.arena.load_sgd_state_dict(opt_sd)
                elif self.optimizer is not None:
                    self.optimizer.load_state_dict(opt_sd)
        if 'plot_stuff' in checkpoint:
            (self.all_tr_losses, self.all_val_losses, self.all_val_losses_tr_mode,
             self.all_val_eval_metrics) = (list(x) for x in checkpoint['plot_stuff'])
            # :683-693 (old off-by-one of finished runs); a checkpoint without a loss history keeps its epoch
            if self.all_tr_losses and self.epoch != len(self.all_tr_losses):
                self.epoch = len(self.all_tr_losses)
                self.all_tr_losses = self.all_tr_losses[:self.epoch]
                self.all_val_losses = self.all_val_losses[:self.epoch]
                self.all_val_losses_tr_mode = self.all_val_losses_tr_mode[:self.epoch]
                self.all_val_eval_metrics = self.all_val_eval_metrics[:self.epoch]
        if 'best_stuff' in checkpoint:
            (self.best_epoch_based_on_MA_tr_loss, self.best_MA_tr_loss_for_patience,
             self.best_val_eval_criterion_MA) = checkpoint['best_stuff']
        if train:
            self.maybe_update_lr(self.epoch)


class MultiTalent_trainer_ddp_2000ep(MultiTalent_trainer_ddp):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.max_num_epochs = 2000


# BASELINE.json names the trainer `nnUNetTrainerV2_MultiTalent`; the reference class is `MultiTalent_trainer_ddp`.
nnUNetTrainerV2_MultiTalent = MultiTalent_trainer_ddp
