"""TEST INFRASTRUCTURE -- makes the UNMODIFIED reference (`/root/reference`) importable in the build container.

This file is part of `oracle/`: it may only be used by `tests/`, by `oracle/make_golden.py` (the committed
fixture generator) and by nothing on the product path.  `/root/reference` does not exist on the GPU box, so nothing
here may be imported from a `-m gpu` test, `smoke()` or `bench.py`.

Recipe: put the reference on `sys.path`; serve empty stub packages for its missing third-party
roots (batchgenerators, SimpleITK, nibabel, skimage, medpy, matplotlib, ...); give `batchgenerators` the two real
helpers the hot path calls (`join/isfile/...` and `pad_nd_image`, the latter taken from our own restatement in
`oracle/unet_oracle.py::pad_nd_image`); set the three nnU-Net path env vars to scratch dirs; start a gloo world of
size 1 so `awesome_allgather_function` (nnunet/utilities/distributed.py:28-73) works on CPU.  No reference source
is copied or modified.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("MTB200_REFERENCE_ROOT", "/root/reference")

_MISSING_ROOTS = (
    "batchgenerators", "SimpleITK", "nibabel", "skimage", "medpy", "matplotlib", "dicom2nifti", "tifffile",
    "unittest2", "monai", "seaborn", "IPython",
)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nnunet"))


class _Dummy:
    """Stand-in for any class imported from a stubbed third-party module (only ever subclassed / referenced)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None


class _StubModule(types.ModuleType):
    __all__ = []
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _MISSING_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _decorate(module)


def _decorate(module):
    """Real implementations for the handful of stubbed helpers the hot path actually executes."""
    name = module.__name__
    if name == "batchgenerators.utilities.file_and_folder_operations":
        import pickle
        import json

        def maybe_mkdir_p(d):
            os.makedirs(d, exist_ok=True)

        def load_pickle(f, mode="rb"):
            with open(f, mode) as fh:
                return pickle.load(fh)

        def save_pickle(obj, f, mode="wb"):
            with open(f, mode) as fh:
                pickle.dump(obj, fh)

        def subfiles(folder, join=True, prefix=None, suffix=None, sort=True):
            res = [os.path.join(folder, i) if join else i for i in os.listdir(folder)
                   if os.path.isfile(os.path.join(folder, i)) and (prefix is None or i.startswith(prefix))
                   and (suffix is None or i.endswith(suffix))]
            return sorted(res) if sort else res

        ns = dict(os=os, join=os.path.join, isfile=os.path.isfile, isdir=os.path.isdir, maybe_mkdir_p=maybe_mkdir_p,
                  load_pickle=load_pickle, save_pickle=save_pickle, subfiles=subfiles, json=json, pickle=pickle)
        module.__dict__.update(ns)
        module.__all__ = list(ns)
    elif name == "batchgenerators.augmentations.utils":
        from oracle.unet_oracle import pad_nd_image  # our restatement of the un-vendored third-party helper
        module.pad_nd_image = pad_nd_image
    elif name == "batchgenerators.dataloading.data_loader":
        # restated base class of the un-vendored batchgenerators (>= 0.23): its constructor only stores its arguments
        class SlimDataLoaderBase(object):
            def __init__(self, data, batch_size, number_of_threads_in_multithreaded=None):
                self._data, self.batch_size = data, batch_size
                self.number_of_threads_in_multithreaded = number_of_threads_in_multithreaded
                self.thread_id = 0

            def __iter__(self):
                return self

            def __next__(self):
                return self.generate_train_batch()
        module.SlimDataLoaderBase = SlimDataLoaderBase


_installed = False


def install():
    """Idempotent.  After this, `import nnunet...` resolves to the unmodified reference."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s (expected only in the build container)" % REFERENCE_ROOT)
    scratch = tempfile.mkdtemp(prefix="mtb200_ref_")
    for k in ("nnUNet_raw_data_base", "nnUNet_preprocessed", "RESULTS_FOLDER"):
        os.environ.setdefault(k, os.path.join(scratch, k))
        os.makedirs(os.environ[k], exist_ok=True)
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    _installed = True


def init_gloo_single():
    """World of size 1 on gloo so the reference's differentiable all-gather runs on CPU."""
    import torch.distributed as dist
    if dist.is_available() and not dist.is_initialized():
        f = tempfile.NamedTemporaryFile(prefix="mtb200_gloo_", delete=False)
        f.close()
        os.unlink(f.name)
        dist.init_process_group("gloo", init_method="file://" + f.name, rank=0, world_size=1)


def build_reference_generic_unet(num_input_channels=1, base_num_features=30, num_classes=47, pool_op_kernel_sizes=None,
                                 conv_kernel_sizes=None, conv_per_stage=2, seed=0):
    """Build the reference network exactly as nnUNetTrainerV2.initialize_network does
    (nnunet/training/network_training/nnUNetTrainerV2.py:131-164)."""
    install()
    import torch
    from torch import nn
    from nnunet.network_architecture.generic_UNet import Generic_UNet
    from nnunet.network_architecture.initialization import InitWeights_He
    if pool_op_kernel_sizes is None:
        pool_op_kernel_sizes = [[2, 2, 2]] * 4 + [[1, 2, 2]]
    if conv_kernel_sizes is None:
        conv_kernel_sizes = [[3, 3, 3]] * (len(pool_op_kernel_sizes) + 1)
    torch.manual_seed(seed)
    net = Generic_UNet(num_input_channels, base_num_features, num_classes, len(pool_op_kernel_sizes), conv_per_stage, 2,
                       nn.Conv3d, nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d,
                       {'p': 0, 'inplace': True}, nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True}, True, False,
                       lambda x: x, InitWeights_He(1e-2), pool_op_kernel_sizes, conv_kernel_sizes, False, True, True)
    net.inference_apply_nonlin = nn.Sigmoid()  # MultiTalent_Trainer_DDP.py:46
    return net


def reference_compute_loss(output, target, valid_regions, ds_loss_weights):
    """Call the reference's unbound MultiTalent_trainer_ddp.compute_loss (MultiTalent_Trainer_DDP.py:544-623)."""
    install()
    init_gloo_single()
    from types import SimpleNamespace
    from torch import nn
    from nnunet.training.network_training.custom_trainers.MultiTalent.MultiTalent.MultiTalent_Trainer_DDP import \
        MultiTalent_trainer_ddp
    ns = SimpleNamespace(ce_loss=nn.BCEWithLogitsLoss(), batch_dice=True, ds_loss_weights=ds_loss_weights)
    return MultiTalent_trainer_ddp.compute_loss(ns, output, target, valid_regions)
