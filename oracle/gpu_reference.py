"""TEST / BASELINE INFRASTRUCTURE -- the reference's hot path on the GPU through the LIBRARY (cuDNN / ATen), i.e. what the
unmodified reference does on a B200: `torch.autocast` around the network + loss, `GradScaler`, `clip_grad_norm_(12)`,
`torch.optim.SGD(nesterov)` (MultiTalent_Trainer_DDP.py:324-370, nnUNetTrainerV2.py:166-170), with
`cudnn.benchmark = True` (network_trainer.py:60-69).  It is the oracle port (`oracle/unet_oracle.py`, pinned against the
reference) executed on CUDA tensors: the functional torch ops dispatch to the same cuDNN / ATen kernels the reference's
`nn.Module`s call.  `/root/reference` does not exist on the GPU box, hence the port.

Used twice, never on the product path:
  * tests (`-m gpu`): T1 parity of the tcgen05 path against the reference under autocast on the same GPU, and
    independent per-layer checks of every tensor-core kernel against cuDNN fp32 on identically rounded operands;
  * `bench.py`: the `gpu_reference` leg -- "the real bar" of SURVEY.md section 2a / 8(d): patches/s of the library path
    next to the hand-written kernels.
"""
import contextlib

import torch
import torch.nn.functional as F

from . import unet_oracle as O


@contextlib.contextmanager
def strict_fp32():
    """fp32 library reference without TF32 (cuDNN convs default to allow_tf32=True)."""
    c, m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = c, m


class GpuReference:
    """Generic_UNet (or FabiansUNet) + MultiTalent loss + optimizer step as the reference runs them on a GPU.

    `sd`: reference-format state_dict (any device); `amp_dtype`: None (fp32) | torch.float16 (the reference's
    `fp16=True` default: `autocast()`) | torch.bfloat16; `channels_last`: run the convolutions on channels_last_3d
    tensors (the reference does not; offered so the library arm can be shown at its best)."""

    def __init__(self, sd, pool, convk, amp_dtype=None, channels_last=False, lr=1e-2, device="cuda", arch="generic",
                 blocks_enc=None, blocks_dec=None, ds_loss_weights=None):
        self.pool, self.convk, self.arch = pool, convk, arch
        self.blocks_enc, self.blocks_dec = blocks_enc, blocks_dec
        self.amp_dtype, self.channels_last = amp_dtype, channels_last
        self.names = list(sd.keys())
        self.params = {}
        for k, v in sd.items():
            p = v.detach().to(device=device, dtype=torch.float32).clone()
            if channels_last and p.dim() == 5:
                p = p.contiguous(memory_format=torch.channels_last_3d)
            self.params[k] = p.requires_grad_(True)
        # nnUNetTrainerV2.py:166-170
        self.optimizer = torch.optim.SGD(list(self.params.values()), lr, weight_decay=3e-5, momentum=0.99, nesterov=True)
        # network_trainer.py:400-402 (GradScaler whenever fp16); a bf16 autocast run keeps it (it is a no-op numerically)
        self.scaler = torch.amp.GradScaler("cuda", enabled=amp_dtype is not None)
        n = len(pool)
        self.w = ds_loss_weights if ds_loss_weights is not None else O.multitalent_ds_loss_weights(n)
        torch.backends.cudnn.benchmark = True

    def _autocast(self):
        if self.amp_dtype is None:
            return contextlib.nullcontext()
        return torch.autocast("cuda", dtype=self.amp_dtype)

    def _net(self, x, do_ds=True):
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last_3d)
        if self.arch == "generic":
            return O.generic_unet_forward(x, self.params, self.pool, self.convk, do_ds=do_ds)
        return O.fabians_unet_forward(x, self.params, self.blocks_enc, self.pool, self.convk, self.blocks_dec,
                                      do_ds=do_ds)

    def forward(self, x, do_ds=True):
        with torch.no_grad(), self._autocast():
            return self._net(x, do_ds)

    def forward_loss(self, x, targets, valid_regions):
        with self._autocast():
            out = self._net(x)
            return out, O.multitalent_loss(out, targets, valid_regions, self.w)

    def train_step(self, x, targets, valid_regions):
        """MultiTalent_Trainer_DDP.py:340-355 (fp16 branch) / :356-363 (fp32 branch)."""
        self.optimizer.zero_grad()
        _, (l, ce, dc) = self.forward_loss(x, targets, valid_regions)
        if self.amp_dtype is not None:
            self.scaler.scale(l).backward()
            self.scaler.unscale_(self.optimizer)
            torch.nn.utils.clip_grad_norm_(list(self.params.values()), 12)
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            l.backward()
            torch.nn.utils.clip_grad_norm_(list(self.params.values()), 12)
            self.optimizer.step()
        return l.detach(), ce.detach(), dc.detach()

    def grads(self, x, targets, valid_regions):
        """(loss triple, {name: unscaled fp32 gradient}) of one forward/backward, no optimizer step."""
        for p in self.params.values():
            p.grad = None
        _, (l, ce, dc) = self.forward_loss(x, targets, valid_regions)
        l.backward()
        return (l.detach(), ce.detach(), dc.detach()), {k: p.grad.detach().clone() for k, p in self.params.items()}


def conv_reference(x_ncdhw, w, stride, padding, transposed=False):
    """cuDNN fp32 (no TF32) conv of operands that were already rounded to the 16-bit storage type: the independent
    reference for one tensor-core kernel launch (products of 16-bit values are exact in fp32; only the summation order
    differs)."""
    with strict_fp32():
        if transposed:
            return F.conv_transpose3d(x_ncdhw.float(), w.float(), None, stride=stride)
        return F.conv3d(x_ncdhw.float(), w.float(), None, stride=stride, padding=padding)


def conv_reference_grads(x_ncdhw, w, dy, stride, padding, transposed=False):
    """(dgrad, wgrad) of the above by autograd through the library op, fp32 without TF32."""
    x = x_ncdhw.float().detach().requires_grad_(True)
    wt = w.float().detach().requires_grad_(True)
    with strict_fp32():
        y = F.conv_transpose3d(x, wt, None, stride=stride) if transposed else F.conv3d(x, wt, None, stride=stride,
                                                                                      padding=padding)
        y.backward(dy.float())
    return x.grad, wt.grad


def mask_dice(prob_a, prob_b, thr=0.5):
    """Dice agreement of the thresholded masks of two probability volumes (all channels pooled)."""
    a, b = prob_a > thr, prob_b > thr
    inter = (a & b).sum().double()
    den = a.sum().double() + b.sum().double()
    return float((2 * inter / den.clamp(min=1)).item()) if float(den) > 0 else 1.0
