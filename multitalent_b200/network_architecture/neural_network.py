"""`SegmentationNetwork` -- mirror of nnunet/network_architecture/neural_network.py:48-163 (predict_3D) and :245-428,
:502-591 (Gaussian map, step list, tiled sliding window, mirror TTA) with the aggregation moved into HBM.

Same method names, argument meaning, return contract (numpy seg [X,Y,Z] + numpy probabilities [C,X,Y,Z]) and error
behaviour as the reference; what differs is where the work happens:

 * tiles are gathered on the device straight from the resident volume (flip folded into the gather indexing),
 * every (mirrored) forward's logits go through ONE fused kernel: sigmoid * 1/n_mirror * gaussian, un-flip, scatter-add
   into fp32 accumulators [C, X, Y, Z] that never leave HBM (the reference does `.cpu().numpy()` of 739 MB per tile
   and a host numpy `+=`, neural_network.py:391-394),
 * ONE single-channel weight volume replaces the reference's C identical copies (`aggregated_nb_of_predictions`,
   :363,372,394) -- same quotient,
 * normalisation + in-order threshold (:405, :415-417) is one kernel; a single D2H of the result at the end.
"""
from typing import List, Tuple, Union

import os
import numpy as np
import torch
from torch import nn

import ctypes
from ctypes import c_void_p as C_void

from .. import _lib as L
from ..engine import Feat, ndhwc_view_info


def _to_host_numpy(*tensors):
    """Device tensors -> numpy arrays through page-locked staging buffers (torch's caching host allocator keeps them for
    the next volume): the [47, X, Y, Z] fp32 probability volume is 6-25 GB, and a pageable `.cpu()` moves it at a few
    GB/s.  Falls back to the pageable copy if page-locked memory cannot be had."""
    outs, pinned = [], []
    try:
        for t in tensors:
            h = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            h.copy_(t, non_blocking=True)
            pinned.append(h)
        torch.cuda.current_stream().synchronize()
        outs = [h.numpy() for h in pinned]
    except RuntimeError:
        outs = [t.cpu().numpy() for t in tensors]
    return outs


class NeuralNetwork(nn.Module):
    def __init__(self):
        super(NeuralNetwork, self).__init__()

    def get_device(self):
        p = next(self.parameters())
        return "cpu" if p.device.type == "cpu" else p.device.index

    def set_device(self, device):
        if device == "cpu":
            self.cpu()
        else:
            self.cuda(device)

    def forward(self, x):
        raise NotImplementedError


def pad_nd_image(image, new_shape=None, mode="constant", kwargs=None, return_slicer=False,
                 shape_must_be_divisible_by=None):
    """Host-side symmetric padding up to `new_shape` (batchgenerators.augmentations.utils.pad_nd_image, third party;
    call site neural_network.py:301): floor(diff/2) below, rest above; returns the slicer that undoes it."""
    if kwargs is None:
        kwargs = {'constant_values': 0}
    if new_shape is None:
        assert shape_must_be_divisible_by is not None
        new_shape = image.shape[-len(shape_must_be_divisible_by):]
    nd = len(new_shape)
    cur = np.array(image.shape[-nd:])
    want = np.maximum(np.array(new_shape), cur)
    if shape_must_be_divisible_by is not None:
        div = np.array(shape_must_be_divisible_by if isinstance(shape_must_be_divisible_by, (list, tuple, np.ndarray))
                       else [shape_must_be_divisible_by] * nd)
        want = (want + div - 1) // div * div
    extra = want - cur
    lo = extra // 2
    hi = extra - lo
    pads = [[0, 0]] * (image.ndim - nd) + [[int(a), int(b)] for a, b in zip(lo, hi)]
    res = np.pad(image, pads, mode, **kwargs) if extra.any() else image
    if not return_slicer:
        return res
    return res, [slice(p[0], res.shape[i] - p[1]) for i, p in enumerate(pads)]


_MIRROR_FLIPS = [(), (4,), (3,), (4, 3), (2,), (4, 2), (3, 2), (4, 3, 2)]  # order of neural_network.py:531-586


def _flip_bits(dims):
    """tensor dims (2,3,4) = (D,H,W) -> flip bitmask used by the kernels (bit0 W, bit1 H, bit2 D)."""
    return sum({4: 1, 3: 2, 2: 4}[d] for d in dims)


class SegmentationNetwork(NeuralNetwork):
    def __init__(self):
        super(NeuralNetwork, self).__init__()
        self.input_shape_must_be_divisible_by = None
        self.conv_op = None
        self.num_classes = None
        self.inference_apply_nonlin = lambda x: x
        self._gaussian_3d = self._patch_size_for_gaussian_3d = None
        self._gaussian_2d = self._patch_size_for_gaussian_2d = None
        self._gaussian_3d_dev = None

    # ---- hooks a concrete network provides for the native predictor ------------------------------------------------
    def native_logits(self, tile: Feat) -> Feat:
        """Full-resolution logits (NDHWC) for an NDHWC input tile, no tape.  Implemented by Generic_UNet/FabiansUNet."""
        raise NotImplementedError

    def native_input_channels_padded(self) -> int:
        raise NotImplementedError

    def native_dtype(self):
        raise NotImplementedError

    def _nonlin_mode(self) -> int:
        """The aggregation kernel applies the inference non-linearity itself: 1 = sigmoid (MultiTalent, MT:43-46),
        2 = softmax over the channels (`softmax_helper` of the single-task / fine-tuning trainers, nnUNetTrainerV2.py:162).
        Anything else must not be aggregated silently with the wrong function."""
        f = self.inference_apply_nonlin
        if isinstance(f, nn.Sigmoid):
            return 1
        if (isinstance(f, nn.Softmax) and f.dim == 1) or getattr(f, "__name__", "") == "softmax_helper":
            return 2
        raise NotImplementedError("the native sliding-window predictor fuses the inference non-linearity into its "
                                  "aggregation kernel (sigmoid or channel softmax); inference_apply_nonlin=%r is not "
                                  "supported" % (f,))

    # ---- reference API ----------------------------------------------------------------------------------------------
    def predict_3D(self, x: np.ndarray, do_mirroring: bool, mirror_axes: Tuple[int, ...] = (0, 1, 2),
                   use_sliding_window: bool = False, step_size: float = 0.5, patch_size: Tuple[int, ...] = None,
                   regions_class_order: Tuple[int, ...] = None, use_gaussian: bool = False,
                   pad_border_mode: str = "constant", pad_kwargs: dict = None, all_in_gpu: bool = False,
                   verbose: bool = True, mixed_precision: bool = True, region_vec=None,
                   return_device_tensors: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """neural_network.py:73-163.  `x` is (c, x, y, z) numpy; returns (segmentation, class probabilities).
        `all_in_gpu` / `mixed_precision` are accepted for signature compatibility: accumulators are always fp32 in HBM
        and the arithmetic type is the network's native dtype.  `return_device_tensors=True` (extension) skips the
        final D2H copy and returns CUDA tensors."""
        assert step_size <= 1, 'step_size must be smaller than 1. Otherwise there will be a gap between consecutive ' \
                               'predictions'
        if pad_kwargs is None:
            pad_kwargs = {'constant_values': 0}
        if len(mirror_axes):
            if self.conv_op == nn.Conv2d and max(mirror_axes) > 1:
                raise ValueError("mirror axes. duh")
            if self.conv_op == nn.Conv3d and max(mirror_axes) > 2:
                raise ValueError("mirror axes. duh")
        if self.training:
            print('WARNING! Network is in train mode during inference. This may be intended, or not...')
        assert len(x.shape) == 4, "data must have shape (c,x,y,z)"
        if self.conv_op != nn.Conv3d:
            raise RuntimeError("the native predictor implements the 3D-conv path only (3d_fullres)")
        if region_vec is not None:
            raise NotImplementedError("region_vec conditioning is not part of the MultiTalent 3d_fullres path")
        self._nonlin_mode()
        with torch.no_grad():
            if use_sliding_window:
                return self._internal_predict_3D_3Dconv_tiled(x, step_size, do_mirroring, mirror_axes, patch_size,
                                                              regions_class_order, use_gaussian, pad_border_mode,
                                                              pad_kwargs, all_in_gpu, verbose,
                                                              return_device_tensors=return_device_tensors)
            return self._internal_predict_3D_3Dconv(x, patch_size, do_mirroring, mirror_axes, regions_class_order,
                                                    pad_border_mode, pad_kwargs, verbose,
                                                    return_device_tensors=return_device_tensors)

    @staticmethod
    def _get_gaussian(patch_size, sigma_scale=1. / 8) -> np.ndarray:
        """neural_network.py:245-259 (scipy's gaussian_filter is host-side and runs once per patch size)."""
        from scipy.ndimage import gaussian_filter
        tmp = np.zeros(patch_size)
        tmp[tuple(i // 2 for i in patch_size)] = 1
        g = gaussian_filter(tmp, [i * sigma_scale for i in patch_size], 0, mode='constant', cval=0)
        g = (g / np.max(g) * 1).astype(np.float32)
        g[g == 0] = np.min(g[g != 0])  # zero weights would produce NaNs in the normalisation
        return g

    @staticmethod
    def _compute_steps_for_sliding_window(patch_size: Tuple[int, ...], image_size: Tuple[int, ...],
                                          step_size: float) -> List[List[int]]:
        """neural_network.py:261-285."""
        assert [i >= j for i, j in zip(image_size, patch_size)], "image size must be as large or larger than patch_size"
        assert 0 < step_size <= 1, 'step_size must be larger than 0 and smaller or equal to 1'
        out = []
        for p, im in zip(patch_size, image_size):
            n = int(np.ceil((im - p) / (p * step_size))) + 1
            span = im - p
            actual = span / (n - 1) if n > 1 else 99999999999
            out.append([int(np.round(actual * i)) for i in range(n)])
        return out

    # ---- native tiled predictor ---------------------------------------------------------------------------------
    def _mirror_list(self, do_mirroring, mirror_axes):
        if not do_mirroring:
            return [()], 1
        keep = [dims for dims in _MIRROR_FLIPS if all((d - 2) in mirror_axes for d in dims)]
        return keep, 2 ** len(mirror_axes)

    def _internal_predict_3D_3Dconv_tiled(self, x: np.ndarray, step_size: float, do_mirroring: bool, mirror_axes: tuple,
                                          patch_size: tuple, regions_class_order: tuple, use_gaussian: bool,
                                          pad_border_mode: str, pad_kwargs: dict, all_in_gpu: bool, verbose: bool,
                                          region_vec=None, return_device_tensors=False):
        """neural_network.py:287-428."""
        assert len(x.shape) == 4, "x must be (c, x, y, z)"
        assert patch_size is not None, "patch_size cannot be None for tiled prediction"
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise L.Mtb200Error("the native predictor needs the network on a CUDA device (no CPU fallback)")
        patch_size = tuple(int(p) for p in patch_size)
        data, slicer = pad_nd_image(x, patch_size, pad_border_mode, pad_kwargs, True, None)
        Cin, X, Y, Z = data.shape
        steps = self._compute_steps_for_sliding_window(patch_size, data.shape[1:], step_size)
        num_tiles = len(steps[0]) * len(steps[1]) * len(steps[2])
        if verbose:
            print("data shape:", data.shape, "patch size:", patch_size, "steps:", steps, "number of tiles:", num_tiles)

        gauss = None
        if use_gaussian and num_tiles > 1:
            if self._gaussian_3d is None or tuple(self._patch_size_for_gaussian_3d) != patch_size:
                self._gaussian_3d = self._get_gaussian(patch_size, sigma_scale=1. / 8)
                self._patch_size_for_gaussian_3d = patch_size
                self._gaussian_3d_dev = None
            if self._gaussian_3d_dev is None or self._gaussian_3d_dev.device != dev:
                self._gaussian_3d_dev = torch.from_numpy(self._gaussian_3d).to(dev)
            gauss = self._gaussian_3d_dev

        vol = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(dev, non_blocking=True)
        C = self.num_classes
        acc = torch.zeros((C, X, Y, Z), dtype=torch.float32, device=dev)
        nb = torch.zeros((X, Y, Z), dtype=torch.float32, device=dev)
        mirrors, n_results = self._mirror_list(do_mirroring, mirror_axes)
        nonlin = self._nonlin_mode()
        dt = self.native_dtype()
        cin_p = self.native_input_channels_padded()
        pd, ph, pw = patch_size
        # TB tiles go through the network as one batch (the deep levels of a single 192x160x128 tile cannot fill 148 SMs)
        TB = max(1, int(os.environ.get("MTB200_INFER_TB", 0)) or int(getattr(self, "inference_tile_batch", 8)))
        tile = torch.empty((TB, pd, ph, pw, cin_p), dtype=dt, device=dev)
        st = L.stream_ptr()
        work = [(sx, sy, sz, mi, dims) for sx in steps[0] for sy in steps[1] for sz in steps[2]
                for mi, dims in enumerate(mirrors)]
        tile_elems = pd * ph * pw * cin_p
        esize = tile.element_size()
        # Results to the host WHILE the remaining tiles are computed: the tiles come in ascending x order, so every plane
        # below the next tile's origin is final -- it is normalised, thresholded and copied (page-locked buffers, own
        # stream) as soon as the batch that completed it has been queued.  25.8 GB of fp32 probabilities for a 512^3
        # volume are ~0.5 s of PCIe time, about two thirds of the compute time of the same volume.
        order = None
        if regions_class_order is not None:
            order = torch.tensor([float(c) for c in regions_class_order], dtype=torch.float32, device=dev)
            assert order.numel() == C
        stream_out = None
        if not return_device_tensors and getattr(self, "stream_results_to_host", True) and num_tiles > 1:
            try:
                crop = [(s.start or 0, n if s.stop is None else s.stop) for s, n in zip(slicer[1:], (X, Y, Z))]
                (xs, xe), (ys, ye), (zs, ze) = crop
                stream_out = {"prob": torch.empty((C, xe - xs, ye - ys, ze - zs), dtype=torch.float32, pin_memory=True),
                              "seg": torch.empty((xe - xs, ye - ys, ze - zs), dtype=torch.float32, pin_memory=True),
                              "dseg": torch.empty((X, Y, Z), dtype=torch.float32, device=dev),
                              "stream": torch.cuda.Stream(dev), "done": 0, "crop": crop}
            except RuntimeError:
                stream_out = None  # no page-locked memory: one copy at the end

        def flush_slab(x1):
            so = stream_out
            x0 = so["done"]
            if x1 <= x0:
                return
            so["done"] = x1
            off = x0 * Y * Z * 4
            L.call("mtb200_sw_finalize_slab", C_void(acc.data_ptr() + off), C_void(nb.data_ptr() + off), C, X * Y * Z,
                   (x1 - x0) * Y * Z, L.ptr(order), C_void(so["dseg"].data_ptr() + off), st)
            (xs, xe), (ys, ye), (zs, ze) = so["crop"]
            c0, c1 = max(x0, xs), min(x1, xe)
            if c0 >= c1:
                return
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            full_yz = ys == 0 and ye == Y and zs == 0 and ze == Z
            with torch.cuda.stream(so["stream"]):
                so["stream"].wait_event(ev)
                for c in range(C):
                    src = acc[c, c0:c1] if full_yz else acc[c, c0:c1, ys:ye, zs:ze].contiguous()
                    so["prob"][c, c0 - xs:c1 - xs].copy_(src, non_blocking=True)
                src = so["dseg"][c0:c1] if full_yz else so["dseg"][c0:c1, ys:ye, zs:ze].contiguous()
                so["seg"][c0 - xs:c1 - xs].copy_(src, non_blocking=True)

        for w0 in range(0, len(work), TB):
            if stream_out is not None and w0 > 0:
                flush_slab(work[w0][0])  # nothing from here on touches the planes below this tile's origin
            chunk = work[w0:w0 + TB]
            for b, (sx, sy, sz, mi, dims) in enumerate(chunk):
                L.call("mtb200_sw_gather_tile", L.ptr(vol), Cin, X, Y, Z, sx, sy, sz, pd, ph, pw, _flip_bits(dims),
                       C_void(tile.data_ptr() + b * tile_elems * esize), L.dtype_enum(dt), cin_p, st)
            eng = getattr(self, "_engine", None)
            fuse = eng is not None and getattr(eng, "fuse_head_aggregate", False) and dt != torch.float32
            if fuse:  # ask the network for the head's input: head -> non-linearity x Gaussian -> scatter-add is ONE kernel
                eng.capture_head, eng.captured_head = True, None
            try:
                logits = self.native_logits(Feat(tile[:len(chunk)], 0, Cin, cin_p))
            finally:
                if fuse:
                    eng.capture_head = False
            cap = eng.captured_head if fuse else None
            if cap is not None:
                hx, hop = cap
                eng.captured_head = None
                assert hop.Cout == C and tuple(hx.dims[1:]) == (pd, ph, pw)
                hp = L.HeadAggParams()
                hp.w_fwd = hop.packed(eng.wdtype, False).data_ptr()
                bias = None
                if hop.bias is not None:
                    bias = torch.zeros(hop.Cout_p, dtype=torch.float32, device=dev)
                    bias[:hop.Cout] = hop.bias.detach().float()
                hp.bias = bias.data_ptr() if bias is not None else None
                hp.gauss = gauss.data_ptr() if gauss is not None else None
                hp.acc, hp.weight = acc.data_ptr(), 1.0 / n_results
                hp.dtype, hp.x_ldc, hp.x_coff, hp.Cin, hp.Cout, hp.C = L.dtype_enum(dt), hx.ldc, hx.coff, hop.Cin_p, hop.Cout_p, C
                hp.pd, hp.ph, hp.pw, hp.nonlin = pd, ph, pw, nonlin
                hp.X, hp.Y, hp.Z = X, Y, Z
                xstride = pd * ph * pw * hx.ldc * esize
                for b, (sx, sy, sz, mi, dims) in enumerate(chunk):
                    hp.x = hx.ptr() + b * xstride
                    hp.nb = nb.data_ptr() if mi == 0 else None
                    hp.flip, hp.x0, hp.y0, hp.z0 = _flip_bits(dims), sx, sy, sz
                    L.call("mtb200_head_aggregate", ctypes.byref(hp), st)
                continue
            lstride = pd * ph * pw * logits.ldc * esize
            for b, (sx, sy, sz, mi, dims) in enumerate(chunk):
                L.call("mtb200_sw_aggregate", C_void(logits.ptr() + b * lstride), L.dtype_enum(dt), logits.ldc, C, pd, ph,
                       pw, _flip_bits(dims), L.ptr(gauss), 1.0 / n_results, nonlin, L.ptr(acc), L.ptr(nb) if mi == 0 else None,
                       X, Y, Z, sx, sy, sz, st)
        if stream_out is not None:
            flush_slab(X)
            stream_out["stream"].synchronize()
            seg_np, acc_np = stream_out["seg"].numpy(), stream_out["prob"].numpy()
            if regions_class_order is None:
                seg_np = seg_np.astype(np.int64)
            if verbose:
                print("prediction done")
            return seg_np, acc_np
        # undo the padding (neural_network.py:397-402) -- crop BEFORE normalising, as the reference does
        sl = tuple([slice(0, C)] + list(slicer[1:]))
        if any(s.start != 0 or s.stop != n for s, n in zip(sl[1:], (X, Y, Z))):
            acc = acc[sl].contiguous()
            nb = nb[tuple(slicer[1:])].contiguous()
        nvox = nb.numel()
        seg = torch.empty(nb.shape, dtype=torch.float32, device=dev)
        L.call("mtb200_sw_finalize", L.ptr(acc), L.ptr(nb), C, nvox, L.ptr(order), L.ptr(seg), st)
        if regions_class_order is None:
            seg = seg.long()
        if return_device_tensors:
            return seg, acc
        if verbose:
            print("prediction done")
        seg_np, acc_np = _to_host_numpy(seg, acc)
        return seg_np, acc_np

    def _internal_predict_3D_3Dconv(self, x: np.ndarray, min_size: Tuple[int, ...], do_mirroring: bool,
                                    mirror_axes: tuple = (0, 1, 2), regions_class_order: tuple = None,
                                    pad_border_mode: str = "constant", pad_kwargs: dict = None, verbose: bool = True,
                                    region_vec=None, return_device_tensors=False):
        """Fully convolutional variant (neural_network.py:465-500): pad to divisibility, one (mirrored) pass."""
        assert len(x.shape) == 4, "x must be (c, x, y, z)"
        assert self.input_shape_must_be_divisible_by is not None
        data, slicer = pad_nd_image(x, min_size, pad_border_mode, pad_kwargs, True,
                                    self.input_shape_must_be_divisible_by)
        res = self._internal_predict_3D_3Dconv_tiled(data, 1.0, do_mirroring, mirror_axes, tuple(data.shape[1:]),
                                                     regions_class_order, False, pad_border_mode, pad_kwargs, False,
                                                     False, return_device_tensors=True)
        seg, prob = res
        sl = tuple(slicer[1:])
        seg, prob = seg[sl], prob[(slice(None),) + sl]
        if return_device_tensors:
            return seg, prob
        seg_np, prob_np = _to_host_numpy(seg.contiguous(), prob.contiguous())
        return seg_np, prob_np

    def _internal_maybe_mirror_and_pred_3D(self, x: Union[np.ndarray, torch.Tensor], mirror_axes: tuple,
                                           do_mirroring: bool = True, mult: Union[np.ndarray, torch.Tensor] = None,
                                           region_vec=None) -> torch.Tensor:
        """neural_network.py:502-591: (1/n) sum over mirrors of un-flipped sigmoid(net(flipped x)), times `mult`.
        Returns a CUDA fp32 tensor [1, C, X, Y, Z] like the reference."""
        assert len(x.shape) == 5, 'x must be (b, c, x, y, z)'
        assert x.shape[0] == 1, "the reference calls this with one tile at a time"
        nonlin = self._nonlin_mode()
        dev = next(self.parameters()).device
        xt = torch.as_tensor(x, dtype=torch.float32, device=dev)[0].contiguous()
        Cin, X, Y, Z = xt.shape
        C = self.num_classes
        acc = torch.zeros((C, X, Y, Z), dtype=torch.float32, device=dev)
        g = None if mult is None else torch.as_tensor(mult, dtype=torch.float32, device=dev).contiguous()
        mirrors, n_results = self._mirror_list(do_mirroring, mirror_axes)
        dt = self.native_dtype()
        cin_p = self.native_input_channels_padded()
        tile = torch.empty((1, X, Y, Z, cin_p), dtype=dt, device=dev)
        st = L.stream_ptr()
        for dims in mirrors:
            fb = _flip_bits(dims)
            L.call("mtb200_sw_gather_tile", L.ptr(xt), Cin, X, Y, Z, 0, 0, 0, X, Y, Z, fb, L.ptr(tile), L.dtype_enum(dt),
                   cin_p, st)
            logits = self.native_logits(Feat(tile, 0, Cin, cin_p))
            L.call("mtb200_sw_aggregate", logits.ptr(), L.dtype_enum(dt), logits.ldc, C, X, Y, Z, fb, L.ptr(g),
                   1.0 / n_results, nonlin, L.ptr(acc), None, X, Y, Z, 0, 0, 0, st)
        return acc[None]
