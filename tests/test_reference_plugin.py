"""The drop-in boundary exercised with the LIVE reference: the shipped plugin file is resolved by
the reference's own by-name lookup, its classes satisfy the reference's `issubclass(nnUNetTrainer)` assertion, and their
hooks build the native network with the reference's checkpoint keys.  Build container only (/root/reference)."""
import os
import shutil
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir("/root/reference/nnunet")
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container)")


@pytest.fixture(scope="module")
def plugin_dir(tmp_path_factory):
    from oracle import ref_import
    ref_import.install()
    d = tmp_path_factory.mktemp("plug")
    pkg = d / "mtb_plugins"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    shutil.copy(os.path.join(ROOT, "multitalent_b200", "integration", "MultiTalent_trainer_b200.py"), str(pkg))
    sys.path.insert(0, str(d))
    yield str(pkg)
    sys.path.remove(str(d))


@pytest.mark.parametrize("name,base", [("MultiTalent_trainer_ddp_b200", "MultiTalent_trainer_ddp"),
                                       ("MultiTalent_trainer_resenc_ddp_b200", "MultiTalent_trainer_resenc_ddp"),
                                       ("nnUNetTrainerV2_MultiTalent", "MultiTalent_trainer_ddp")])
def test_reference_lookup_finds_the_plugin_classes(plugin_dir, name, base):
    from nnunet.training.model_restore import recursive_find_python_class
    from nnunet.training.network_training.nnUNetTrainer import nnUNetTrainer
    cls = recursive_find_python_class([plugin_dir], name, current_module="mtb_plugins")
    assert cls is not None, "the reference's by-name lookup (model_restore.py:23-41) did not find %s" % name
    assert issubclass(cls, nnUNetTrainer)                       # run_training_DDP.py:158-159, model_restore.py:78
    assert base in [c.__name__ for c in cls.__mro__]


def test_plugin_hooks_build_the_native_network_with_reference_keys(plugin_dir):
    from nnunet.network_architecture.generic_UNet import Generic_UNet as RefNet
    from nnunet.network_architecture.initialization import InitWeights_He
    from nnunet.network_architecture.neural_network import SegmentationNetwork as RefSeg
    from nnunet.training.model_restore import recursive_find_python_class
    from torch import nn
    cls = recursive_find_python_class([plugin_dir], "MultiTalent_trainer_ddp_b200", current_module="mtb_plugins")
    tr = object.__new__(cls)            # the reference __init__ starts NCCL (nnUNetTrainerV2_DDP.py:68): not on the CPU
    tr.num_input_channels, tr.base_num_features, tr.num_classes, tr.conv_per_stage, tr.fp16 = 1, 8, 47, 2, True
    tr.net_num_pool_op_kernel_sizes = [[2, 2, 2], [2, 2, 2], [1, 2, 2]]
    tr.net_conv_kernel_sizes = [[3, 3, 3]] * 4
    tr.initialize_network()
    net = tr.network
    assert isinstance(net, RefSeg)                              # nnUNetTrainerV2_DDP.py:617-619
    assert net._native_ok and net.native_dtype() == torch.bfloat16 and isinstance(net.inference_apply_nonlin, nn.Sigmoid)
    ref = RefNet(1, 8, 47, 3, 2, 2, nn.Conv3d, nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d,
                 {'p': 0, 'inplace': True}, nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True}, True, False,
                 lambda x: x, InitWeights_He(1e-2), tr.net_num_pool_op_kernel_sizes, tr.net_conv_kernel_sizes, False, True,
                 True)
    a, b = ref.state_dict(), net.state_dict()
    assert list(a.keys()) == list(b.keys()) and all(a[k].shape == b[k].shape for k in a)
    net.load_state_dict(a)                                      # a reference checkpoint loads into the native network
    # the loss hook refuses CPU tensors loudly (no fallback) rather than computing something else
    from multitalent_b200._lib import Mtb200Error
    with pytest.raises(Mtb200Error):
        tr.ds_loss_weights = [1.0]
        tr.compute_loss((torch.zeros(1, 47, 4, 4, 4),), (torch.zeros(1, 1, 4, 4, 4),), [("03_liver",)])


def test_resenc_plugin_builds_fabians_unet(plugin_dir):
    from nnunet.training.model_restore import recursive_find_python_class
    from multitalent_b200.plans import default_plans
    cls = recursive_find_python_class([plugin_dir], "MultiTalent_trainer_resenc_ddp_b200", current_module="mtb_plugins")
    tr = object.__new__(cls)
    tr.plans, tr.stage = default_plans("resenc"), 1
    tr.num_input_channels, tr.base_num_features, tr.num_classes, tr.fp16 = 1, 30, 47, False
    tr.initialize_network()
    assert tr.network._native_ok and sum(p.numel() for p in tr.network.parameters()) == 69335475
    assert all(float(b.norm2.weight.abs().sum()) == 0.0 for st in tr.network.encoder.stages for b in st.convs)
