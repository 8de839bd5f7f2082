"""`MultiTalent_trainer_resenc_ddp` -- the hot-path half of
nnunet/training/network_training/custom_trainers/MultiTalent/MultiTalent/MultiTalent_meets_resenc.py:36-216, 713-798 on
the native kernels: residual-encoder U-Net (`FabiansUNet`, plan keys `num_blocks_encoder` / `num_blocks_decoder`), the
same multi-head BCE + pooled-Dice loss as `MultiTalent_trainer_ddp`, deep-supervision scales taken from
`pool_op_kernel_sizes[1:]` (:108-116) and weights 1/2^i with the lowest output unweighted (:158-170).
"""
import numpy as np
import torch
from torch import nn

from ...network_architecture.generic_UNet import InitWeights_He
from ...network_architecture.generic_modular_residual_UNet import (FabiansUNet, get_default_network_config,
                                                                    init_last_bn_before_add_to_0)
from .MultiTalent_Trainer_DDP import MultiTalent_trainer_ddp


class MultiTalent_trainer_resenc_ddp(MultiTalent_trainer_ddp):
    def initialize_network(self):
        """MultiTalent_meets_resenc.py:72-103."""
        assert self.threeD, "MultiTalent is a 3d_fullres configuration"
        cfg = get_default_network_config(3, None, norm_type="in")
        sp = self.plans['plans_per_stage'][self.stage]
        self.network = FabiansUNet(self.num_input_channels, self.base_num_features, sp['num_blocks_encoder'], 2,
                                   sp['pool_op_kernel_sizes'], sp['conv_kernel_sizes'], cfg, self.num_classes,
                                   sp['num_blocks_decoder'], True, False, 320, InitWeights_He(1e-2),
                                   native_dtype=self.native_dtype)
        self.network.apply(init_last_bn_before_add_to_0)
        if torch.cuda.is_available():
            self.network.cuda()
        self.network.inference_apply_nonlin = nn.Sigmoid()

    def setup_DA_params(self):
        """:108-116 -- the first pooling entry belongs to the first (unstrided) stage and is skipped."""
        self.deep_supervision_scales = [[1, 1, 1]] + list(
            list(i) for i in 1 / np.cumprod(np.vstack(self.net_num_pool_op_kernel_sizes[1:]), axis=0))[:-1]


class MultiTalent_trainer_resenc_ddp_2000ep(MultiTalent_trainer_resenc_ddp):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.max_num_epochs = 2000
