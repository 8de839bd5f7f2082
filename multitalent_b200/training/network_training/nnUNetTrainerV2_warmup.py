"""`nnUNetTrainerV2_warmupsegheads` -- the fine-tuning trainer of the reference's second workflow (readme.md:51-66;
nnunet/training/network_training/nnUNet_variants/pretraining/nnUNetTrainerV2_warmup.py:67-199) on the native kernels:

  * epochs [0, warmup_duration): only the segmentation heads train, AdamW(3e-3, amsgrad) with the learning rate rising
    linearly to `warmup_max_lr` (:89-95, :122-124);
  * epochs [warmup_duration, +num_epochs_sgd_warmup): the whole network, Nesterov SGD, learning rate rising linearly to
    `initial_lr` (:96-100, :116-119);
  * afterwards: poly learning rate over the remaining epochs (:101-112).

The network is the same `Generic_UNet` as in the MultiTalent trainers (softmax inference non-linearity), the loss the
softmax Dice + CE with deep supervision (nnUNetTrainer.py:108, nnUNetTrainerV2.py:77-90) on `mtb200_dcce_*`.  The
reference computes every gradient during the heads-only phase and lets the optimizer ignore the trunk's; here
`freeze_trunk_during_head_warmup` (default on) marks the trunk parameters as not requiring a gradient for that phase, which
puts `Generic_UNet` on its frozen-trunk path (backward = the heads' weight gradients only).  The heads receive identical
gradients either way BEFORE clipping (`tests/test_gpu_warmup_trainer.py`).  Known deviation: the reference's
`clip_grad_norm_(network.parameters(), 12)` also counts the (unused) trunk gradients, so whenever the full-network
gradient norm exceeds 12 its head gradients are scaled by a smaller coefficient than here; reproducing that coefficient
needs the trunk gradients, i.e. the full backward -- pass `freeze_trunk_during_head_warmup=False` for exact reference
behaviour.

Data loading, augmentation, validation and checkpoint files stay in the reference."""
import numpy as np
import torch
from torch import nn

from ...network_architecture.generic_UNet import Generic_UNet, InitWeights_He
from ...plans import default_plans
from ...run.load_pretrained_weights import load_pretrained_weights
from ..loss_functions.deep_supervision import MultipleOutputLoss2
from ..loss_functions.dice_loss import DC_and_CE_loss
from .MultiTalent_Trainer_DDP import poly_lr


def softmax_helper(x):
    """nnunet/utilities/nd_softmax.py:20."""
    return torch.softmax(x, 1)


def warmup_lr(epoch, warmup_duration, num_epochs_sgd_warmup, warmup_max_lr, initial_lr, max_num_epochs):
    """Learning rate the reference sets at the start of `epoch` (nnUNetTrainerV2_warmup.py:87-112, called with
    epoch=None so that `self.epoch` is used)."""
    if epoch < warmup_duration:
        return (epoch + 1) / warmup_duration * warmup_max_lr
    if epoch < warmup_duration + num_epochs_sgd_warmup:
        return (epoch - warmup_duration + 1) / num_epochs_sgd_warmup * initial_lr
    ep = epoch - (warmup_duration + num_epochs_sgd_warmup - 1)
    assert ep > 0, "epoch must be >0"
    return poly_lr(ep, max_num_epochs - num_epochs_sgd_warmup - warmup_duration, initial_lr, 0.9)


class nnUNetTrainerV2_warmupsegheads(object):
    def __init__(self, plans_file, fold, output_folder=None, dataset_directory=None, batch_dice=True, stage=None,
                 unpack_data=True, deterministic=True, fp16=False, native_dtype=None,
                 freeze_trunk_during_head_warmup=True):
        """Positional signature of nnUNetTrainerV2_warmup.py:68-71.  Keyword-only extensions: `native_dtype`,
        `freeze_trunk_during_head_warmup`."""
        self.init_args = (plans_file, fold, output_folder, dataset_directory, batch_dice, stage, unpack_data,
                          deterministic, fp16)
        self.plans_file, self.fold = plans_file, fold
        self.output_folder, self.dataset_directory = output_folder, dataset_directory
        self.batch_dice, self.stage, self.fp16 = batch_dice, stage, fp16
        self.native_dtype = native_dtype if native_dtype is not None else (torch.float16 if fp16 else torch.float32)
        self.freeze_trunk_during_head_warmup = freeze_trunk_during_head_warmup
        self.initial_lr, self.weight_decay = 1e-2, 3e-5       # nnUNetTrainerV2.py:47-48
        self.num_epochs_sgd_warmup = 50                       # :73
        self.warmup_max_lr = 5e-4                             # :74
        self.warmup_duration = 10                             # :75
        self.max_num_epochs = 1000 + self.num_epochs_sgd_warmup + self.warmup_duration  # :76
        self.epoch = 0
        self.lr = None
        self.was_initialized = False
        self.plans = self.network = self.optimizer = self.loss = None
        self.lr_scheduler = None

    # ---- plans / initialisation --------------------------------------------------------------------------------------
    def process_plans(self, plans):
        """nnUNetTrainer.py:326-392, the fields the hot path needs (num_classes = foreground classes + background)."""
        if self.stage is None:
            self.stage = max(plans['plans_per_stage'].keys())
        sp = plans['plans_per_stage'][self.stage]
        self.plans = plans
        self.batch_size = int(sp['batch_size'])
        self.patch_size = np.array(sp['patch_size']).astype(int)
        self.net_num_pool_op_kernel_sizes = [list(map(int, k)) for k in sp['pool_op_kernel_sizes']]
        self.net_conv_kernel_sizes = [list(map(int, k)) for k in sp['conv_kernel_sizes']]
        self.base_num_features = int(plans['base_num_features'])
        self.num_input_channels = int(plans['num_modalities'])
        self.conv_per_stage = int(plans.get('conv_per_stage', 2))
        self.num_classes = int(plans['num_classes']) + 1

    def initialize(self, training=True, force_load_plans=False):
        if self.was_initialized:
            return
        if isinstance(self.plans_file, dict):
            plans = self.plans_file
        elif self.plans_file is None:
            plans = default_plans()
        else:
            import pickle
            with open(self.plans_file, 'rb') as f:
                plans = pickle.load(f)
        self.process_plans(plans)
        self.deep_supervision_scales = [[1, 1, 1]] + list(
            list(i) for i in 1 / np.cumprod(np.vstack(self.net_num_pool_op_kernel_sizes), axis=0))[:-1]
        n = len(self.net_num_pool_op_kernel_sizes)
        w = np.array([1 / (2 ** i) for i in range(n)])
        w[n - 1] = 0                                          # nnUNetTrainerV2.py:84-87
        self.ds_loss_weights = w / w.sum()
        self.loss = MultipleOutputLoss2(
            DC_and_CE_loss({'batch_dice': self.batch_dice, 'smooth': 1e-5, 'do_bg': False}, {}), self.ds_loss_weights)
        self.initialize_network()
        self.was_initialized = True
        if training:
            self.initialize_optimizer_and_scheduler(True)     # heads first (:81-85)
            self.maybe_update_lr()

    def initialize_network(self):
        """nnUNetTrainerV2.py:131-164."""
        self.network = Generic_UNet(self.num_input_channels, self.base_num_features, self.num_classes,
                                    len(self.net_num_pool_op_kernel_sizes), self.conv_per_stage, 2, nn.Conv3d,
                                    nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d,
                                    {'p': 0, 'inplace': True}, nn.LeakyReLU,
                                    {'negative_slope': 1e-2, 'inplace': True}, True, False, lambda x: x,
                                    InitWeights_He(1e-2), self.net_num_pool_op_kernel_sizes,
                                    self.net_conv_kernel_sizes, False, True, True, native_dtype=self.native_dtype)
        if torch.cuda.is_available():
            self.network.cuda()
        self.network.inference_apply_nonlin = softmax_helper

    def load_pretrained_weights(self, fname, verbose=False):
        """run_training.py:186-188 (`-pretrained_weights`)."""
        if not self.was_initialized:
            self.initialize(True)
        return load_pretrained_weights(self.network, fname, verbose)

    def initialize_optimizer_and_scheduler(self, seg_heads_only=False):
        """nnUNetTrainerV2_warmup.py:122-134."""
        assert self.network is not None, "self.initialize_network must be called first"
        trunk = [p for n, p in self.network.named_parameters() if not n.startswith(self.head_prefix)]
        if seg_heads_only:
            self.optimizer = torch.optim.AdamW(self.head_module().parameters(), 3e-3,
                                               weight_decay=self.weight_decay, amsgrad=True)
            if self.freeze_trunk_during_head_warmup:
                for p in trunk:
                    p.requires_grad_(False)
                    p.grad = None
        else:
            for p in trunk:
                p.requires_grad_(True)
            self.optimizer = torch.optim.SGD(self.network.parameters(), self.initial_lr,
                                             weight_decay=self.weight_decay, momentum=0.99, nesterov=True)
        self.seg_heads_only = seg_heads_only
        self.lr_scheduler = None
        # network_trainer.py:400-402: GradScaler whenever the arithmetic is fp16 (dynamic loss scaling)
        if getattr(self, "amp_grad_scaler", None) is None:
            self.amp_grad_scaler = (torch.amp.GradScaler("cuda")
                                    if self.native_dtype == torch.float16 and torch.cuda.is_available() else None)

    head_prefix = "seg_outputs."

    def head_module(self):
        """The segmentation heads the first phase trains (nnUNetTrainerV2_warmup.py:123)."""
        return self.network.seg_outputs

    def maybe_update_lr(self, epoch=None):
        """:87-112 (the reference's training loop calls it without an argument: `self.epoch` decides)."""
        ep = self.epoch if epoch is None else epoch
        lr = warmup_lr(ep, self.warmup_duration, self.num_epochs_sgd_warmup, self.warmup_max_lr, self.initial_lr,
                       self.max_num_epochs)
        self.optimizer.param_groups[0]['lr'] = lr
        self.lr = lr
        return lr

    def on_epoch_end(self):
        """:114-120 -- switch to whole-network SGD when the head warm-up is over, then network_trainer.on_epoch_end
        (network_trainer.py:603-616): `maybe_update_lr()` runs with `self.epoch` still naming the epoch that just ended,
        and only afterwards does the training loop increment the counter (:482-490).  So epoch e >= 1 trains at
        warmup_lr(e - 1) and epoch `warmup_duration` is still a heads-only AdamW epoch -- exactly the reference's
        sequence (tests/test_warmup_trainer_host.py drives this loop against the reference's)."""
        if self.epoch == self.warmup_duration:
            self.initialize_optimizer_and_scheduler(seg_heads_only=False)
        self.maybe_update_lr()
        self.epoch += 1
        return self.epoch < self.max_num_epochs

    # ---- the hot path ------------------------------------------------------------------------------------------------
    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False):
        """nnUNetTrainerV2.py:219-271: fetch, forward, DC+CE with deep supervision, backward, clip 12, step; returns the
        loss as a numpy scalar."""
        data_dict = next(data_generator)
        data = torch.as_tensor(data_dict['data'])
        target = [torch.as_tensor(t) for t in data_dict['target']]
        if torch.cuda.is_available():
            data = data.cuda(non_blocking=True)
            target = [t.cuda(non_blocking=True) for t in target]
        self.optimizer.zero_grad()
        with torch.set_grad_enabled(do_backprop):
            output = self.network(data)
            l = self.loss(output, target)
        if do_backprop:
            params = [p for p in self.network.parameters() if p.requires_grad]
            if self.amp_grad_scaler is not None:   # nnUNetTrainerV2.py:246-252
                self.amp_grad_scaler.scale(l).backward()
                self.amp_grad_scaler.unscale_(self.optimizer)
                torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 12)
                self.amp_grad_scaler.step(self.optimizer)
                self.amp_grad_scaler.update()
            else:
                l.backward()
                torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 12)
                self.optimizer.step()
        return l.detach().cpu().numpy()

    def predict_preprocessed_data_return_seg_and_softmax(self, data, do_mirroring=True, mirror_axes=None,
                                                         use_sliding_window=True, step_size=0.5, use_gaussian=True,
                                                         pad_border_mode='constant', pad_kwargs=None, all_in_gpu=False,
                                                         verbose=True, mixed_precision=True):
        """nnUNetTrainerV2.py:198-217: deep supervision off, eval mode, tiled prediction with the softmax inference
        non-linearity fused into the aggregation kernel (`mtb200_sw_aggregate`, nonlin = 2), argmax segmentation."""
        if pad_border_mode == 'constant' and pad_kwargs is None:
            pad_kwargs = {'constant_values': 0}
        if do_mirroring and mirror_axes is None:
            mirror_axes = (0, 1, 2)
        net = self.network
        ds, mode = self._get_ds(), net.training
        self._set_ds(False)
        net.eval()
        try:
            return net.predict_3D(data, do_mirroring=do_mirroring, mirror_axes=mirror_axes or (),
                                  use_sliding_window=use_sliding_window, step_size=step_size,
                                  patch_size=tuple(self.patch_size), regions_class_order=None,
                                  use_gaussian=use_gaussian, pad_border_mode=pad_border_mode, pad_kwargs=pad_kwargs,
                                  all_in_gpu=all_in_gpu, verbose=verbose, mixed_precision=mixed_precision)
        finally:
            net.train(mode)
            self._set_ds(ds)

    def _get_ds(self):
        return self.network.do_ds

    def _set_ds(self, v):
        self.network.do_ds = v


class nnUNetTrainerV2_warmupsegheads_resenc(nnUNetTrainerV2_warmupsegheads):
    """nnUNetTrainerV2_warmup.py:441-560: the same three-phase schedule on the residual-encoder network (`FabiansUNet`
    built from the plan's `num_blocks_encoder` / `num_blocks_decoder`, softmax inference, `init_last_bn_before_add_to_0`);
    the heads are `decoder.deep_supervision_outputs`; deep-supervision scales skip the first (unstrided) pooling entry
    (:483-489); prediction toggles `decoder.deep_supervision` (:507-528)."""
    head_prefix = "decoder.deep_supervision_outputs."

    def head_module(self):
        return self.network.decoder.deep_supervision_outputs

    def initialize(self, training=True, force_load_plans=False):
        if self.was_initialized:
            return
        super().initialize(training, force_load_plans)
        # :483-489 -- recompute scales / loss weights for the resenc pooling list (first entry = unstrided stem stage)
        self.deep_supervision_scales = [[1, 1, 1]] + list(
            list(i) for i in 1 / np.cumprod(np.vstack(self.net_num_pool_op_kernel_sizes[1:]), axis=0))[:-1]

    def initialize_network(self):
        """:451-468."""
        from ...network_architecture.generic_modular_residual_UNet import (FabiansUNet, get_default_network_config,
                                                                            init_last_bn_before_add_to_0)
        cfg = get_default_network_config(3, None, norm_type="in")
        sp = self.plans['plans_per_stage'][self.stage]
        self.network = FabiansUNet(self.num_input_channels, self.base_num_features, sp['num_blocks_encoder'], 2,
                                   sp['pool_op_kernel_sizes'], sp['conv_kernel_sizes'], cfg, self.num_classes,
                                   sp['num_blocks_decoder'], True, False, 320, InitWeights_He(1e-2),
                                   native_dtype=self.native_dtype)
        if torch.cuda.is_available():
            self.network.cuda()
        self.network.inference_apply_nonlin = softmax_helper
        self.network.apply(init_last_bn_before_add_to_0)

    def _get_ds(self):
        return self.network.decoder.deep_supervision

    def _set_ds(self, v):
        self.network.decoder.deep_supervision = v
