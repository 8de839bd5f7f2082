"""Reference-side plugin file -- copy (or symlink) it into
`nnunet/training/network_training/custom_trainers/MultiTalent/` of a MIC-DKFZ/MultiTalent checkout with `multitalent_b200`
on PYTHONPATH.  The reference resolves trainers BY CLASS NAME through `recursive_find_python_class`
(nnunet/training/model_restore.py:23-41) and requires `issubclass(trainer, nnUNetTrainer)`
(nnunet/run/run_training_DDP.py:158-159, model_restore.py:78): the classes below ARE subclasses of the reference's own
trainers, so everything around the hot path (data loading, augmentation, epoch loop, logging, checkpoints, validation
export) stays the reference's code, and only the network and the loss are replaced by the sm_100a kernels.

    python -m torch.distributed.launch --nproc_per_node=8 nnunet/run/run_training_DDP.py 3d_fullres \
        MultiTalent_trainer_ddp_b200 100 0 -p MultiTalent_bs4 --dbs

Requires the reference (`nnunet`) to be importable; importing this module without it raises ImportError.
"""
import torch
from torch import nn

from nnunet.network_architecture.neural_network import SegmentationNetwork as _RefSegmentationNetwork
from nnunet.training.network_training.custom_trainers.MultiTalent.MultiTalent.MultiTalent_Trainer_DDP import \
    MultiTalent_trainer_ddp
from nnunet.training.network_training.custom_trainers.MultiTalent.MultiTalent.MultiTalent_meets_resenc import \
    MultiTalent_trainer_resenc_ddp

from multitalent_b200.network_architecture import generic_UNet as _g
from multitalent_b200.network_architecture import generic_modular_residual_UNet as _r
from multitalent_b200.training.loss_functions.multitalent_loss import multitalent_loss


# The reference asserts `isinstance(self.network, (SegmentationNetwork, nn.DataParallel, DDP))` before predicting
# (nnUNetTrainerV2_DDP.py:617-619): the native networks additionally derive from the REFERENCE's base class (all methods
# resolve to the native ones first).
class Generic_UNet(_g.Generic_UNet, _RefSegmentationNetwork):
    pass


class FabiansUNet(_r.FabiansUNet, _RefSegmentationNetwork):
    pass


def _native_dtype(trainer):
    # `fp16=True` is the reference's autocast switch: bf16 storage needs no GradScaler interplay with the reference's
    # own `amp_grad_scaler` (it stays a numerical no-op); fp32 otherwise.  The native networks ignore torch.autocast.
    return torch.bfloat16 if trainer.fp16 else torch.float32


class MultiTalent_trainer_ddp_b200(MultiTalent_trainer_ddp):
    def initialize_network(self):                       # replaces nnUNetTrainerV2.py:131-164 + MT:43-46
        self.network = Generic_UNet(self.num_input_channels, self.base_num_features, self.num_classes,
                                    len(self.net_num_pool_op_kernel_sizes), self.conv_per_stage, 2, nn.Conv3d,
                                    nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d,
                                    {'p': 0, 'inplace': True}, nn.LeakyReLU,
                                    {'negative_slope': 1e-2, 'inplace': True}, True, False, lambda x: x,
                                    _g.InitWeights_He(1e-2), self.net_num_pool_op_kernel_sizes,
                                    self.net_conv_kernel_sizes, False, True, True, native_dtype=_native_dtype(self))
        if torch.cuda.is_available():
            self.network.cuda()
        self.network.inference_apply_nonlin = nn.Sigmoid()

    def compute_loss(self, output, target, valid_regions):   # replaces MT:544-623
        return multitalent_loss(output, target, valid_regions, self.ds_loss_weights,
                                engine=getattr(self.network, "_engine", None))


class MultiTalent_trainer_resenc_ddp_b200(MultiTalent_trainer_resenc_ddp):
    def initialize_network(self):                       # replaces MultiTalent_meets_resenc.py:72-104
        cfg = _r.get_default_network_config(3, None, norm_type="in")
        sp = self.plans['plans_per_stage'][self.stage]
        self.network = FabiansUNet(self.num_input_channels, self.base_num_features, sp['num_blocks_encoder'], 2,
                                   sp['pool_op_kernel_sizes'], sp['conv_kernel_sizes'], cfg, self.num_classes,
                                   sp['num_blocks_decoder'], True, False, 320, _g.InitWeights_He(1e-2),
                                   native_dtype=_native_dtype(self))
        self.network.apply(_r.init_last_bn_before_add_to_0)
        if torch.cuda.is_available():
            self.network.cuda()
        self.network.inference_apply_nonlin = nn.Sigmoid()

    def compute_loss(self, output, target, valid_regions):   # replaces MultiTalent_meets_resenc.py:713-798
        return multitalent_loss(output, target, valid_regions, self.ds_loss_weights,
                                engine=getattr(self.network, "_engine", None))


# BASELINE.json's name for the MultiTalent trainer
nnUNetTrainerV2_MultiTalent = MultiTalent_trainer_ddp_b200
