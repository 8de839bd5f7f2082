"""`load_pretrained_weights` -- nnunet/run/load_pretrained_weights.py:17-61, the first half of the fine-tuning workflow
(readme.md:51-66): copy every tensor of a MultiTalent checkpoint whose key AND shape match the target network.  The
segmentation heads of a downstream task have another number of classes, so they never match and keep their
initialisation ("THIS DOES NOT TRANSFER SEGMENTATION HEADS", :19).  Same contract and error behaviour as the reference:
a leading `module.` (DDP / DataParallel checkpoints) is stripped, every `conv_blocks*` tensor of the network must be
present with the same shape or a RuntimeError is raised and nothing is modified.

Difference in mechanism only: tensors are copied IN PLACE (`param.copy_`) instead of `load_state_dict` on a rebuilt dict,
so parameters that are views of the trainer's flat arena stay views."""
from collections import OrderedDict

import torch


def load_pretrained_weights(network, fname, verbose=False):
    """`fname`: path of a `model_final_checkpoint.model`-style file, or an already loaded checkpoint dict."""
    saved_model = fname if isinstance(fname, dict) else torch.load(fname, map_location="cpu")
    pretrained = OrderedDict()
    for k, v in saved_model['state_dict'].items():
        pretrained[k[7:] if k.startswith('module.') else k] = v
    model_dict = network.state_dict()
    for key, value in model_dict.items():
        if 'conv_blocks' in key and not (key in pretrained and tuple(value.shape) == tuple(pretrained[key].shape)):
            raise RuntimeError("Pretrained weights are not compatible with the current network architecture")
    overlap = [k for k, v in pretrained.items() if k in model_dict and tuple(model_dict[k].shape) == tuple(v.shape)]
    print("################### Loading pretrained weights from file ", fname if not isinstance(fname, dict) else "<dict>",
          '###################')
    if verbose:
        print("Below is the list of overlapping blocks in pretrained model and nnUNet architecture:")
        for key in overlap:
            print(key)
    with torch.no_grad():
        for key in overlap:
            model_dict[key].copy_(pretrained[key])
    print("################### Done ###################")
    return overlap
