"""`FabiansUNet` (residual encoder + plain conv decoder) -- drop-in for
nnunet/network_architecture/generic_modular_residual_UNet.py:28-118, 320-358,
nnunet/network_architecture/generic_modular_UNet.py:31-78, 185-291 and
nnunet/network_architecture/custom_modules/conv_blocks.py:49-213, 330-357, whose forward/backward run on the sm_100a
kernels of libmtb200.so (the `MultiTalent_resenc_bs4` configuration of MultiTalent_meets_resenc.py:62-116).

As for `Generic_UNet`, the module tree (attribute names, parameter shapes, registration order = `state_dict` keys:
`encoder.initial_conv.weight`, `encoder.stages.S.convs.B.{conv1,norm1,conv2,norm2,downsample_skip.{0,1}}.*`,
`decoder.tus.I.weight`, `decoder.stages.I.convs.C.{conv,norm}.*`, `decoder.deep_supervision_outputs.I.*`) is the
reference's, the leaf modules are ordinary torch modules, and configurations outside the native envelope (2D, BatchNorm,
dropout, average-pool skips, bottleneck blocks, upscaled logits) run through them.  The MultiTalent configuration on a
CUDA tensor never takes that route.
"""
from copy import deepcopy

import numpy as np
import torch
from torch import nn

from .. import _lib as L
from ..engine import ConvOp, Engine, Feat, PackGroup, collect_ops, pad_channels
from .generic_UNet import Upsample, _UNetFunction
from .neural_network import SegmentationNetwork


def get_default_network_config(dim=2, dropout_p=None, nonlin="LeakyReLU", norm_type="bn"):
    """generic_modular_UNet.py:31-78: the `props` dictionary (op classes + default kwargs)."""
    if dim not in (2, 3):
        raise NotImplementedError
    props = {'conv_op': nn.Conv2d if dim == 2 else nn.Conv3d,
             'dropout_op': nn.Dropout2d if dim == 2 else nn.Dropout3d}
    if norm_type == "bn":
        props['norm_op'] = nn.BatchNorm2d if dim == 2 else nn.BatchNorm3d
    elif norm_type == "in":
        props['norm_op'] = nn.InstanceNorm2d if dim == 2 else nn.InstanceNorm3d
    else:
        raise NotImplementedError
    props['norm_op_kwargs'] = {'eps': 1e-5, 'affine': True}
    if dropout_p is None:
        props['dropout_op'] = None
        props['dropout_op_kwargs'] = {'p': 0, 'inplace': True}
    else:
        props['dropout_op_kwargs'] = {'p': dropout_p, 'inplace': True}
    props['conv_op_kwargs'] = {'stride': 1, 'dilation': 1, 'bias': False}  # kernel size is set by the network
    if nonlin == "LeakyReLU":
        props['nonlin'], props['nonlin_kwargs'] = nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True}
    elif nonlin == "ReLU":
        props['nonlin'], props['nonlin_kwargs'] = nn.ReLU, {'inplace': True}
    else:
        raise ValueError
    return props


def _as_list(conv_op, v):
    if isinstance(v, (tuple, list, np.ndarray)):
        return [int(i) if i is not None else 1 for i in v]
    return [v] * {nn.Conv1d: 1, nn.Conv2d: 2, nn.Conv3d: 3}[conv_op]


class ConvDropoutNormReLU(nn.Module):
    """conv_blocks.py:49-85 (children `conv`, `do`, `norm`, `nonlin`; `all` is the Sequential over them)."""

    def __init__(self, input_channels, output_channels, kernel_size, network_props):
        super().__init__()
        props = deepcopy(network_props)
        kernel_size = _as_list(props['conv_op'], kernel_size)
        self.conv = props['conv_op'](input_channels, output_channels, kernel_size,
                                     padding=[(i - 1) // 2 for i in kernel_size], **props['conv_op_kwargs'])
        self.do = props['dropout_op'](**props['dropout_op_kwargs']) if props['dropout_op'] is not None else nn.Identity()
        self.norm = props['norm_op'](output_channels, **props['norm_op_kwargs']) if props['norm_op'] is not None \
            else nn.Identity()
        self.nonlin = props['nonlin'](**props['nonlin_kwargs'])
        self.all = nn.Sequential(self.conv, self.do, self.norm, self.nonlin)

    def forward(self, x):
        return self.all(x)


class StackedConvLayers(nn.Module):
    """conv_blocks.py:88-113."""

    def __init__(self, input_channels, output_channels, kernel_size, network_props, num_convs, first_stride=None):
        super().__init__()
        props = deepcopy(network_props)
        first = deepcopy(network_props)
        if first_stride is not None:
            first['conv_op_kwargs']['stride'] = first_stride
        self.convs = nn.Sequential(ConvDropoutNormReLU(input_channels, output_channels, kernel_size, first),
                                   *[ConvDropoutNormReLU(output_channels, output_channels, kernel_size, props)
                                     for _ in range(num_convs - 1)])

    def forward(self, x):
        return self.convs(x)


class BasicResidualBlock(nn.Module):
    """conv_blocks.py:116-213: conv1(stride)-norm1-nonlin1-conv2-norm2-(+skip)-nonlin2; the skip is a strided bias-free
    1x1(x1) conv + norm when the stride or the width changes, else the identity."""

    def __init__(self, in_planes, out_planes, kernel_size, props, stride=None, use_avgpool_in_skip=False):
        super().__init__()
        props = deepcopy(props)
        del props['conv_op_kwargs']['stride']
        conv_op = props['conv_op']
        kernel_size = _as_list(conv_op, kernel_size)
        stride = _as_list(conv_op, stride if stride is not None else 1)
        self.stride, self.kernel_size, self.props = stride, kernel_size, props
        self.in_planes, self.out_planes = in_planes, out_planes
        pad = [(i - 1) // 2 for i in kernel_size]
        self.conv1 = conv_op(in_planes, out_planes, kernel_size=kernel_size, padding=pad, stride=stride,
                             **props['conv_op_kwargs'])
        self.norm1 = props['norm_op'](out_planes, **props['norm_op_kwargs'])
        self.nonlin1 = props['nonlin'](**props['nonlin_kwargs'])
        self.dropout = props['dropout_op'](**props['dropout_op_kwargs']) if props['dropout_op_kwargs']['p'] != 0 \
            else nn.Identity()
        self.conv2 = conv_op(out_planes, out_planes, kernel_size=kernel_size, padding=pad, stride=1,
                             **props['conv_op_kwargs'])
        self.norm2 = props['norm_op'](out_planes, **props['norm_op_kwargs'])
        self.nonlin2 = props['nonlin'](**props['nonlin_kwargs'])
        self.use_avgpool_in_skip = use_avgpool_in_skip
        if any(i != 1 for i in stride) or in_planes != out_planes:
            if use_avgpool_in_skip:
                ops = []
                if any(i != 1 for i in stride):
                    ops.append({nn.Conv1d: nn.AvgPool1d, nn.Conv2d: nn.AvgPool2d, nn.Conv3d: nn.AvgPool3d}[conv_op](
                        stride, stride))
                ops.append(conv_op(in_planes, out_planes, kernel_size=1, padding=0, stride=1, bias=False))
                ops.append(props['norm_op'](out_planes, **props['norm_op_kwargs']))
                self.downsample_skip = nn.Sequential(*ops)
            else:
                self.downsample_skip = nn.Sequential(
                    conv_op(in_planes, out_planes, kernel_size=1, padding=0, stride=stride, bias=False),
                    props['norm_op'](out_planes, **props['norm_op_kwargs']))
        else:
            self.downsample_skip = lambda x: x

    def forward(self, x):
        out = self.nonlin1(self.norm1(self.dropout(self.conv1(x))))
        out = self.norm2(self.conv2(out))
        out = out + self.downsample_skip(x)
        return self.nonlin2(out)


class ResidualLayer(nn.Module):
    """conv_blocks.py:330-357 (BasicResidualBlock stacks; the first block carries the stride)."""

    def __init__(self, input_channels, output_channels, kernel_size, network_props, num_blocks, first_stride=None,
                 block=BasicResidualBlock, block_kwargs=None):
        super().__init__()
        block_kwargs = block_kwargs or {}
        props = deepcopy(network_props)
        self.convs = nn.Sequential(block(input_channels, output_channels, kernel_size, props, first_stride, **block_kwargs),
                                   *[block(output_channels, output_channels, kernel_size, props, **block_kwargs)
                                     for _ in range(num_blocks - 1)])
        self.output_channels = output_channels

    def forward(self, x):
        return self.convs(x)


class ResidualUNetEncoder(nn.Module):
    """generic_modular_residual_UNet.py:28-118 (includes the bottleneck stage)."""

    def __init__(self, input_channels, base_num_features, num_blocks_per_stage, feat_map_mul_on_downscale,
                 pool_op_kernel_sizes, conv_kernel_sizes, props, default_return_skips=True, max_num_features=480,
                 block=BasicResidualBlock, block_kwargs=None):
        super().__init__()
        self.default_return_skips, self.props = default_return_skips, props
        assert len(pool_op_kernel_sizes) == len(conv_kernel_sizes)
        n = len(conv_kernel_sizes)
        if not isinstance(num_blocks_per_stage, (list, tuple)):
            num_blocks_per_stage = [num_blocks_per_stage] * n
        assert len(num_blocks_per_stage) == n
        self.num_blocks_per_stage = num_blocks_per_stage
        self.initial_conv = props['conv_op'](input_channels, base_num_features, 3, padding=1, **props['conv_op_kwargs'])
        self.initial_norm = props['norm_op'](base_num_features, **props['norm_op_kwargs'])
        self.initial_nonlin = props['nonlin'](**props['nonlin_kwargs'])
        stages = []
        self.stage_output_features, self.stage_pool_kernel_size, self.stage_conv_op_kernel_size = [], [], []
        cur = base_num_features
        for s in range(n):
            out_f = min(base_num_features * feat_map_mul_on_downscale ** s, max_num_features)
            st = ResidualLayer(cur, out_f, conv_kernel_sizes[s], props, num_blocks_per_stage[s], pool_op_kernel_sizes[s],
                               block, block_kwargs or {})
            stages.append(st)
            self.stage_output_features.append(st.output_channels)
            self.stage_conv_op_kernel_size.append(conv_kernel_sizes[s])
            self.stage_pool_kernel_size.append(pool_op_kernel_sizes[s])
            cur = st.output_channels
        self.output_features = cur
        self.stages = nn.ModuleList(stages)

    def forward(self, x, return_skips=None):
        skips = []
        x = self.initial_nonlin(self.initial_norm(self.initial_conv(x)))
        for s in self.stages:
            x = s(x)
            if self.default_return_skips:
                skips.append(x)
        if return_skips is None:
            return_skips = self.default_return_skips
        return skips if return_skips else x


class PlainConvUNetDecoder(nn.Module):
    """generic_modular_UNet.py:185-291: transposed conv -> cat -> conv stack -> 1x1x1 head (bias) at every level."""

    def __init__(self, previous, num_classes, num_blocks_per_stage=None, network_props=None, deep_supervision=False,
                 upscale_logits=False):
        super().__init__()
        self.num_classes, self.deep_supervision = num_classes, deep_supervision
        self.props = previous.props if network_props is None else network_props
        conv_op = self.props['conv_op']
        if conv_op == nn.Conv2d:
            transpconv, upsample_mode = nn.ConvTranspose2d, "bilinear"
        elif conv_op == nn.Conv3d:
            transpconv, upsample_mode = nn.ConvTranspose3d, "trilinear"
        else:
            raise ValueError("unknown convolution dimensionality, conv op: %s" % str(conv_op))
        if num_blocks_per_stage is None:
            num_blocks_per_stage = previous.num_blocks_per_stage[:-1][::-1]
        assert len(num_blocks_per_stage) == len(previous.num_blocks_per_stage) - 1
        self.stage_pool_kernel_size = previous.stage_pool_kernel_size
        self.stage_output_features = previous.stage_output_features
        self.stage_conv_op_kernel_size = previous.stage_conv_op_kernel_size
        n = len(previous.stages) - 1
        tus, stages, heads = [], [], []
        cum_upsample = np.cumprod(np.vstack(self.stage_pool_kernel_size), axis=0).astype(int)
        f_skip = None
        for i, s in enumerate(np.arange(n)[::-1]):
            f_below, f_skip = self.stage_output_features[s + 1], self.stage_output_features[s]
            tus.append(transpconv(f_below, f_skip, self.stage_pool_kernel_size[s + 1], self.stage_pool_kernel_size[s + 1],
                                  bias=False))
            stages.append(StackedConvLayers(2 * f_skip, f_skip, self.stage_conv_op_kernel_size[s], self.props,
                                            num_blocks_per_stage[i]))
            if deep_supervision and s != 0:
                seg = conv_op(f_skip, num_classes, 1, 1, 0, 1, 1, bias=True)
                heads.append(nn.Sequential(seg, Upsample(scale_factor=cum_upsample[s], mode=upsample_mode))
                             if upscale_logits else seg)
        heads.append(conv_op(f_skip, num_classes, 1, 1, 0, 1, 1, bias=True))
        self.tus, self.stages = nn.ModuleList(tus), nn.ModuleList(stages)
        self.deep_supervision_outputs = nn.ModuleList(heads)

    def forward(self, skips, gt=None, loss=None):
        skips = skips[::-1]
        x = skips[0]
        outs = []
        for i in range(len(self.tus)):
            x = self.stages[i](torch.cat((self.tus[i](x), skips[i + 1]), dim=1))
            if self.deep_supervision:
                t = self.deep_supervision_outputs[i](x)
                outs.append(loss(t, gt) if gt is not None else t)
            else:
                outs = self.deep_supervision_outputs[i](x)
        return outs[::-1] if self.deep_supervision else outs


class ResidualUNetDecoder(nn.Module):
    """generic_modular_residual_UNet.py:142-270: transposed conv -> cat -> ResidualLayer per level; 1x1x1 heads (bias):
    `deep_supervision_outputs` for every level but the highest-resolution one (only with deep supervision) and
    `segmentation_output` for the highest resolution."""

    def __init__(self, previous, num_classes, num_blocks_per_stage=None, network_props=None, deep_supervision=False,
                 upscale_logits=False, block=BasicResidualBlock, block_kwargs=None):
        super().__init__()
        block_kwargs = block_kwargs or {}
        self.num_classes, self.deep_supervision = num_classes, deep_supervision
        self.props = previous.props if network_props is None else network_props
        conv_op = self.props['conv_op']
        if conv_op == nn.Conv2d:
            transpconv, upsample_mode = nn.ConvTranspose2d, "bilinear"
        elif conv_op == nn.Conv3d:
            transpconv, upsample_mode = nn.ConvTranspose3d, "trilinear"
        else:
            raise ValueError("unknown convolution dimensionality, conv op: %s" % str(conv_op))
        if num_blocks_per_stage is None:
            num_blocks_per_stage = previous.num_blocks_per_stage[:-1][::-1]
        assert len(num_blocks_per_stage) == len(previous.num_blocks_per_stage) - 1
        self.stage_pool_kernel_size = previous.stage_pool_kernel_size
        self.stage_output_features = previous.stage_output_features
        self.stage_conv_op_kernel_size = previous.stage_conv_op_kernel_size
        n = len(previous.stages) - 1
        tus, stages, heads = [], [], []
        cum_upsample = np.cumprod(np.vstack(self.stage_pool_kernel_size), axis=0).astype(int)
        f_skip = None
        for i, s in enumerate(np.arange(n)[::-1]):
            f_below, f_skip = self.stage_output_features[s + 1], self.stage_output_features[s]
            tus.append(transpconv(f_below, f_skip, self.stage_pool_kernel_size[s + 1], self.stage_pool_kernel_size[s + 1],
                                  bias=False))
            stages.append(ResidualLayer(2 * f_skip, f_skip, self.stage_conv_op_kernel_size[s], self.props,
                                        num_blocks_per_stage[i], None, block, block_kwargs))
            if deep_supervision and s != 0:
                seg = conv_op(f_skip, num_classes, 1, 1, 0, 1, 1, bias=True)
                heads.append(nn.Sequential(seg, Upsample(scale_factor=cum_upsample[s], mode=upsample_mode))
                             if upscale_logits else seg)
        self.segmentation_output = conv_op(f_skip, num_classes, 1, 1, 0, 1, 1, bias=True)
        self.tus, self.stages = nn.ModuleList(tus), nn.ModuleList(stages)
        self.deep_supervision_outputs = nn.ModuleList(heads)

    def forward(self, skips):
        skips = skips[::-1]
        x = skips[0]
        outs = []
        for i in range(len(self.tus)):
            x = self.stages[i](torch.cat((self.tus[i](x), skips[i + 1]), dim=1))
            if self.deep_supervision and i != len(self.tus) - 1:
                outs.append(self.deep_supervision_outputs[i](x))
        seg = self.segmentation_output(x)
        if self.deep_supervision:
            outs.append(seg)
            return outs[::-1]
        return seg


def init_last_bn_before_add_to_0(module):
    """MultiTalent_meets_resenc.py:31-34: the second norm of every residual block starts at zero."""
    if isinstance(module, BasicResidualBlock):
        module.norm2.weight = nn.init.constant_(module.norm2.weight, 0)
        module.norm2.bias = nn.init.constant_(module.norm2.bias, 0)


def _native_norm(n):
    return isinstance(n, nn.InstanceNorm3d) and n.affine and not n.track_running_stats and abs(n.eps - 1e-5) < 1e-12


def _native_conv(c, need_k=(1, 3)):
    return (isinstance(c, nn.Conv3d) and tuple(c.dilation) == (1, 1, 1) and c.groups == 1 and c.padding_mode == 'zeros'
            and all(k in need_k for k in c.kernel_size) and tuple(c.padding) == tuple((k - 1) // 2 for k in c.kernel_size)
            and all(s in (1, 2) for s in c.stride))


def _native_lrelu(a):
    return isinstance(a, nn.LeakyReLU) and abs(a.negative_slope - 1e-2) < 1e-12


def _native_block(b):
    """True if a BasicResidualBlock is something the kernels implement exactly."""
    if type(b) is not BasicResidualBlock or b.use_avgpool_in_skip or not isinstance(b.dropout, nn.Identity):
        return False
    if not (_native_conv(b.conv1) and _native_conv(b.conv2) and _native_norm(b.norm1) and _native_norm(b.norm2)
            and _native_lrelu(b.nonlin1) and _native_lrelu(b.nonlin2)):
        return False
    if isinstance(b.downsample_skip, nn.Sequential):
        return _native_conv(b.downsample_skip[0], (1,)) and _native_norm(b.downsample_skip[1])
    return True


def _block_ops(b, split=0):
    """(conv1, norm1, conv2, norm2, (skip conv, skip norm) | None) of one BasicResidualBlock.  `split`: logical
    channels of the first half of a concatenated input (decoder blocks read the concat buffer)."""
    skip = None
    if isinstance(b.downsample_skip, nn.Sequential):
        c = b.downsample_skip[0]
        skip = (ConvOp(c.weight, c.bias, c.kernel_size, c.stride, split=split), b.downsample_skip[1])
    return (ConvOp(b.conv1.weight, b.conv1.bias, b.conv1.kernel_size, b.conv1.stride, split=split), b.norm1,
            ConvOp(b.conv2.weight, b.conv2.bias, b.conv2.kernel_size, b.conv2.stride), b.norm2, skip)


def _run_block(eng, tape, blk, f, out=None):
    """BasicResidualBlock.forward (conv_blocks.py:201-213) on the kernels: conv1-IN-LReLU, conv2-IN, skip (1x1x1 conv-IN
    or the identity), add, LReLU."""
    c1, n1, c2, n2, skip = blk
    a = eng.conv_norm(tape, c1, n1.weight, n1.bias, f)
    a = eng.conv_norm(tape, c2, n2.weight, n2.bias, a, slope=1.0)       # norm2 only: the add comes first
    if skip is not None:
        r = eng.conv_norm(tape, skip[0], skip[1].weight, skip[1].bias, f, slope=1.0)
    else:
        # identity skip: the very tensor conv1 consumed (its materialised activation on the tensor-core path, the raw
        # tensor + pending transform on the norm-on-load path), so that both consumers leave their gradients in the same
        # buffer
        r = f.act if f.act is not None else f
    return eng.residual_act(tape, a, r, out=out)


def _run_encoder(eng, tape, ops, f, dev):
    """ResidualUNetEncoder.forward (generic_modular_residual_UNet.py:98-112) -> list of skips (bottleneck last).  Every
    stage output but the bottleneck is written straight into the second half of the decoder's concat buffer."""
    op, nrm = ops['stem']
    f = eng.conv_norm(tape, op, nrm.weight, nrm.bias, f, need_input_grad=False)
    n_stages = len(ops['enc'])
    skips = []
    for s, blocks in enumerate(ops['enc']):
        for bi, blk in enumerate(blocks):
            out = None
            if bi == len(blocks) - 1 and s < n_stages - 1:
                od = blk[0].out_dims(f.dims)
                cat = eng.new_buf(od, 2 * blk[2].Cout_p, dev)
                out = Feat(cat, blk[2].Cout_p, blk[2].Cout, blk[2].Cout_p)
            f = _run_block(eng, tape, blk, f, out=out)
        skips.append(f)
    return skips


class FabiansUNet(SegmentationNetwork):
    """generic_modular_residual_UNet.py:320-358.  `native_dtype` / `native_impl` are keyword-only extensions."""
    use_this_for_2D_configuration = 1244233721.0
    use_this_for_3D_configuration = 1230348801.0
    default_blocks_per_stage_encoder = (1, 2, 3, 4, 4, 4, 4, 4, 4, 4, 4)
    default_blocks_per_stage_decoder = (1, 1, 1, 1, 1, 1, 1, 1, 1, 1)
    default_min_batch_size = 2

    def __init__(self, input_channels, base_num_features, num_blocks_per_stage_encoder, feat_map_mul_on_downscale,
                 pool_op_kernel_sizes, conv_kernel_sizes, props, num_classes, num_blocks_per_stage_decoder,
                 deep_supervision=False, upscale_logits=False, max_features=512, initializer=None,
                 block=BasicResidualBlock, props_decoder=None, block_kwargs=None, native_dtype=torch.float32,
                 native_impl=0):
        super().__init__()
        self.do_ds = deep_supervision
        self.conv_op = props['conv_op']
        self.num_classes = num_classes
        self.upscale_logits = upscale_logits
        self.encoder = ResidualUNetEncoder(input_channels, base_num_features, num_blocks_per_stage_encoder,
                                           feat_map_mul_on_downscale, pool_op_kernel_sizes, conv_kernel_sizes, props,
                                           default_return_skips=True, max_num_features=max_features, block=block,
                                           block_kwargs=block_kwargs or {})
        props['dropout_op_kwargs']['p'] = 0
        self.decoder = PlainConvUNetDecoder(self.encoder, num_classes, num_blocks_per_stage_decoder,
                                            props if props_decoder is None else props_decoder, deep_supervision,
                                            upscale_logits)
        self.input_shape_must_be_divisible_by = np.prod(np.vstack(pool_op_kernel_sizes), 0, dtype=np.int64)
        if initializer is not None:
            self.apply(initializer)
        self._engine = Engine(native_dtype, native_impl)
        self._ops = None
        self._native_ok = self._check_native()

    # ---- native path -------------------------------------------------------------------------------------------------
    def set_native_dtype(self, dtype, impl=None):
        self._engine = Engine(dtype, self._engine.impl if impl is None else impl)

    def native_dtype(self):
        return self._engine.dtype

    def native_input_channels_padded(self):
        return pad_channels(self.encoder.initial_conv.in_channels)

    def _check_native(self):
        e, d = self.encoder, self.decoder
        if self.conv_op != nn.Conv3d or self.upscale_logits or self.num_classes > 64:
            return False
        if not (_native_conv(e.initial_conv, (3,)) and _native_norm(e.initial_norm) and _native_lrelu(e.initial_nonlin)):
            return False
        for st in e.stages:
            if not all(_native_block(b) for b in st.convs):
                return False
        for t in d.tus:
            if not isinstance(t, nn.ConvTranspose3d) or t.bias is not None or tuple(t.kernel_size) != tuple(t.stride):
                return False
        for st in d.stages:
            for c in st.convs:
                if not (_native_conv(c.conv) and _native_norm(c.norm) and _native_lrelu(c.nonlin)
                        and isinstance(c.do, nn.Identity)):
                    return False
        return all(isinstance(h, nn.Conv3d) and tuple(h.kernel_size) == (1, 1, 1) for h in d.deep_supervision_outputs)

    def _build_ops(self):
        e, d = self.encoder, self.decoder
        ops = {'stem': (ConvOp(e.initial_conv.weight, e.initial_conv.bias, e.initial_conv.kernel_size,
                               e.initial_conv.stride), e.initial_norm)}
        ops['enc'] = [[_block_ops(b) for b in st.convs] for st in e.stages]
        dec = []
        for i in range(len(d.tus)):
            t = d.tus[i]
            tu = ConvOp(t.weight, None, t.kernel_size, t.stride, transposed=True)
            convs = []
            for j, c in enumerate(d.stages[i].convs):
                convs.append((ConvOp(c.conv.weight, c.conv.bias, c.conv.kernel_size, c.conv.stride,
                                     split=t.out_channels if j == 0 else 0), c.norm))
            heads = list(d.deep_supervision_outputs)  # built without deep supervision: only the last level has a head
            h = heads[i] if len(heads) == len(d.tus) else (heads[0] if i == len(d.tus) - 1 else None)
            dec.append((tu, convs, h))
        ops['dec'] = dec
        self._ops = ops
        PackGroup(collect_ops(ops))
        self.__dict__['_ops_key'] = e.initial_conv.weight  # plain attribute: must not be registered as a parameter

    def _native_forward(self, x, tape, only_full_res=False):
        """Kernel sequence of FabiansUNet.forward (generic_modular_residual_UNet.py:350-353): logits Feats, highest
        resolution first."""
        L.lib()  # fail loudly if the CUDA library is missing
        eng = self._engine
        if self._ops is None or self.__dict__.get('_ops_key') is not self.encoder.initial_conv.weight:
            self._build_ops()
        ops = self._ops
        dev = x.buf.device if isinstance(x, Feat) else x.device
        f = x if isinstance(x, Feat) else eng.input_feat(x)
        # frozen-trunk fast path (fine-tuning with only the heads trainable, nnUNetTrainerV2_warmup.py:470-476): nothing
        # below the heads is recorded, the backward pass is the heads' weight gradients alone
        ttape = tape
        if tape is not None and not any(p.requires_grad for n, p in self.named_parameters()
                                        if not n.startswith("decoder.deep_supervision_outputs.")):
            ttape = None
        skips = _run_encoder(eng, ttape, ops, f, dev)
        n_stages = len(skips)
        f = skips[-1]
        logits = []
        nd = len(ops['dec'])
        for i, (tu, convs, head) in enumerate(ops['dec']):
            skip = skips[n_stages - 2 - i]
            cat = skip.buf
            assert tu.Cout_p == skip.Cp and cat.shape[4] == 2 * skip.Cp
            eng.conv_plain(ttape, tu, f, Feat(cat, 0, tu.Cout, tu.Cout_p))
            f = Feat(cat, 0, tu.Cout + skip.C, 2 * skip.Cp)
            for (op, nrm) in convs:
                f = eng.conv_norm(ttape, op, nrm.weight, nrm.bias, f)
            if head is None or (only_full_res and i != nd - 1):
                continue
            logits.append(eng.conv_plain(tape, ConvOpCache.get(self, head), f, need_input_grad=ttape is not None,
                                         head=True))
        return logits[::-1]

    def native_logits(self, tile: Feat) -> Feat:
        """Full-resolution logits for the sliding-window predictor (neural_network.py)."""
        self._engine.begin_step()
        return self._native_forward(tile, None, only_full_res=True)[0]

    def forward(self, x):
        if self._native_ok and x.is_cuda:
            want_ds = bool(self.decoder.deep_supervision)
            params = tuple(self.parameters())
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                n_out = len(self.decoder.tus) if want_ds else 1
                outs = _UNetFunction.apply(self, x, n_out, *params)
            else:
                self._engine.begin_step()
                feats = self._native_forward(x, None, only_full_res=not want_ds)
                outs = tuple(f.as_ncdhw() for f in feats)
            return list(outs) if want_ds else outs[0]
        if self._native_ok and not x.is_cuda:
            raise L.Mtb200Error("FabiansUNet (MultiTalent configuration) runs on the native CUDA path only; got a %s "
                                "tensor. There is no CPU fallback." % x.device)
        return self.decoder(self.encoder(x))


class ResidualUNet(SegmentationNetwork):
    """generic_modular_residual_UNet.py:273-318: residual encoder + RESIDUAL decoder (`ResidualUNetDecoder`) -- the
    `ResidualUNet` module API BASELINE.json's north star names.  Same kernels as `FabiansUNet`; the decoder levels are
    `ResidualLayer`s whose first block reads the concat buffer through both its 3x3x3 conv and its 1x1x1 skip conv.
    `native_dtype` / `native_impl` are keyword-only extensions."""
    use_this_for_batch_size_computation_2D = 858931200.0
    use_this_for_batch_size_computation_3D = 727842816.0
    default_base_num_features = 24
    default_conv_per_stage = (2, 2, 2, 2, 2, 2, 2, 2)

    def __init__(self, input_channels, base_num_features, num_blocks_per_stage_encoder, feat_map_mul_on_downscale,
                 pool_op_kernel_sizes, conv_kernel_sizes, props, num_classes, num_blocks_per_stage_decoder,
                 deep_supervision=False, upscale_logits=False, max_features=512, initializer=None,
                 block=BasicResidualBlock, block_kwargs=None, native_dtype=torch.float32, native_impl=0):
        super().__init__()
        block_kwargs = block_kwargs or {}
        self.do_ds = deep_supervision
        self.conv_op = props['conv_op']
        self.num_classes = num_classes
        self.upscale_logits = upscale_logits
        self.encoder = ResidualUNetEncoder(input_channels, base_num_features, num_blocks_per_stage_encoder,
                                           feat_map_mul_on_downscale, pool_op_kernel_sizes, conv_kernel_sizes, props,
                                           default_return_skips=True, max_num_features=max_features, block=block,
                                           block_kwargs=block_kwargs)
        self.decoder = ResidualUNetDecoder(self.encoder, num_classes, num_blocks_per_stage_decoder, props,
                                           deep_supervision, upscale_logits, block=block, block_kwargs=block_kwargs)
        self.input_shape_must_be_divisible_by = np.prod(np.vstack(pool_op_kernel_sizes), 0, dtype=np.int64)
        if initializer is not None:
            self.apply(initializer)
        self._engine = Engine(native_dtype, native_impl)
        self._ops = None
        self._native_ok = self._check_native()

    def set_native_dtype(self, dtype, impl=None):
        self._engine = Engine(dtype, self._engine.impl if impl is None else impl)

    def native_dtype(self):
        return self._engine.dtype

    def native_input_channels_padded(self):
        return pad_channels(self.encoder.initial_conv.in_channels)

    def _heads(self):
        """1x1x1 head of each decoder level (None where the level has none), lowest resolution first."""
        d = self.decoder
        n = len(d.tus)
        ds = list(d.deep_supervision_outputs)
        return [(ds[i] if i < len(ds) else None) for i in range(n - 1)] + [d.segmentation_output]

    def _check_native(self):
        e, d = self.encoder, self.decoder
        if self.conv_op != nn.Conv3d or self.upscale_logits or self.num_classes > 64:
            return False
        if not (_native_conv(e.initial_conv, (3,)) and _native_norm(e.initial_norm) and _native_lrelu(e.initial_nonlin)):
            return False
        for st in list(e.stages) + list(d.stages):
            if not all(_native_block(b) for b in st.convs):
                return False
        for t in d.tus:
            if not isinstance(t, nn.ConvTranspose3d) or t.bias is not None or tuple(t.kernel_size) != tuple(t.stride):
                return False
        return all(h is None or (isinstance(h, nn.Conv3d) and tuple(h.kernel_size) == (1, 1, 1)) for h in self._heads())

    def _build_ops(self):
        e, d = self.encoder, self.decoder
        ops = {'stem': (ConvOp(e.initial_conv.weight, e.initial_conv.bias, e.initial_conv.kernel_size,
                               e.initial_conv.stride), e.initial_norm),
               'enc': [[_block_ops(b) for b in st.convs] for st in e.stages]}
        dec = []
        for i in range(len(d.tus)):
            t = d.tus[i]
            tu = ConvOp(t.weight, None, t.kernel_size, t.stride, transposed=True)
            blocks = [_block_ops(b, split=t.out_channels if j == 0 else 0) for j, b in enumerate(d.stages[i].convs)]
            dec.append((tu, blocks))
        ops['dec'] = dec
        self._ops = ops
        PackGroup(collect_ops(ops))
        self.__dict__['_ops_key'] = e.initial_conv.weight

    def _native_forward(self, x, tape, only_full_res=False):
        """ResidualUNet.forward (:304-306) = encoder (:98-112) + ResidualUNetDecoder.forward (:224-248)."""
        L.lib()  # fail loudly if the CUDA library is missing
        eng = self._engine
        if self._ops is None or self.__dict__.get('_ops_key') is not self.encoder.initial_conv.weight:
            self._build_ops()
        ops = self._ops
        dev = x.buf.device if isinstance(x, Feat) else x.device
        f = x if isinstance(x, Feat) else eng.input_feat(x)
        skips = _run_encoder(eng, tape, ops, f, dev)
        n_stages = len(skips)
        f = skips[-1]
        heads = self._heads()
        logits = []
        nd = len(ops['dec'])
        for i, (tu, blocks) in enumerate(ops['dec']):
            skip = skips[n_stages - 2 - i]
            cat = skip.buf
            assert tu.Cout_p == skip.Cp and cat.shape[4] == 2 * skip.Cp
            eng.conv_plain(tape, tu, f, Feat(cat, 0, tu.Cout, tu.Cout_p))
            f = Feat(cat, 0, tu.Cout + skip.C, 2 * skip.Cp)
            for blk in blocks:
                f = _run_block(eng, tape, blk, f)
            if heads[i] is None or (only_full_res and i != nd - 1):
                continue
            logits.append(eng.conv_plain(tape, ConvOpCache.get(self, heads[i]), f, head=True))
        return logits[::-1]

    def native_logits(self, tile: Feat) -> Feat:
        self._engine.begin_step()
        return self._native_forward(tile, None, only_full_res=True)[0]

    def forward(self, x):
        if self._native_ok and x.is_cuda:
            want_ds = bool(self.decoder.deep_supervision)
            params = tuple(self.parameters())
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                n_out = len(self.decoder.tus) if want_ds else 1
                outs = _UNetFunction.apply(self, x, n_out, *params)
            else:
                self._engine.begin_step()
                feats = self._native_forward(x, None, only_full_res=not want_ds)
                outs = tuple(f.as_ncdhw() for f in feats)
            return list(outs) if want_ds else outs[0]
        if self._native_ok and not x.is_cuda:
            raise L.Mtb200Error("ResidualUNet (MultiTalent configuration) runs on the native CUDA path only; got a %s "
                                "tensor. There is no CPU fallback." % x.device)
        return self.decoder(self.encoder(x))


class ConvOpCache:
    """ConvOp wrappers of the 1x1x1 heads, keyed by module (heads may be toggled by the deep-supervision flag)."""

    @staticmethod
    def get(net, head):
        cache = net.__dict__.setdefault('_head_ops', {})
        hit = cache.get(id(head))
        if hit is None or hit.weight is not head.weight:
            hit = ConvOp(head.weight, head.bias, head.kernel_size, head.stride)
            cache[id(head)] = hit
        return hit
