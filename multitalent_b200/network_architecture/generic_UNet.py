"""`Generic_UNet` -- drop-in for nnunet/network_architecture/generic_UNet.py:156-442 whose forward/backward run on
hand-written sm_100a kernels.

The module tree (names, parameter shapes, registration order) is the reference's, so `state_dict()` keys, checkpoint
loading (`nnUNetTrainerV2_DDP.load_checkpoint_ram`, :636-669), `load_pretrained_weights` and attribute paths touched by
fine-tuning trainers (`seg_outputs`, `conv_blocks_context[0].blocks[0].conv.weight`) keep working.  The leaf modules are
ordinary torch modules and stay executable by torch: argument combinations outside the native path (2D, BatchNorm,
dropout > 0, max-pool, bilinear upsampling, another `basic_block`) run through them -- i.e. through the very library
ops the reference uses.  The MultiTalent configuration (3D, InstanceNorm, LeakyReLU, strided-conv pooling,
transposed-conv upsampling, `nnUNetTrainerV2.py:156-161`) on a CUDA tensor NEVER takes that route: it requires
libmtb200.so and raises if the library is missing.
"""
from copy import deepcopy

import numpy as np
import torch
from torch import nn

from .. import _lib as L
from ..engine import ConvOp, Engine, Feat, PackGroup, Tape, collect_ops, pad_channels
from .neural_network import SegmentationNetwork


class InitWeights_He(object):
    """nnunet/network_architecture/initialization.py:19-27."""

    def __init__(self, neg_slope=1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)


def softmax_helper(x):
    return torch.softmax(x, 1)


class Upsample(nn.Module):
    """custom_modules/helperModules.py:32-46 (unused on the MultiTalent path; kept for API completeness)."""

    def __init__(self, size=None, scale_factor=None, mode='nearest', align_corners=True):
        super(Upsample, self).__init__()
        self.align_corners, self.mode, self.scale_factor, self.size = align_corners, mode, scale_factor, size

    def forward(self, x):
        return nn.functional.interpolate(x, size=self.size, scale_factor=self.scale_factor, mode=self.mode,
                                         align_corners=self.align_corners)


class ConvDropoutNormNonlin(nn.Module):
    """Parameter container + torch-executable leaf for one conv -> (dropout) -> norm -> nonlin block
    (generic_UNet.py:28-70).  Attribute names `conv`, `instnorm`, `lrelu`, `dropout` are checkpoint keys."""

    def __init__(self, input_channels, output_channels, conv_op=nn.Conv2d, conv_kwargs=None, norm_op=nn.BatchNorm2d,
                 norm_op_kwargs=None, dropout_op=nn.Dropout2d, dropout_op_kwargs=None, nonlin=nn.LeakyReLU,
                 nonlin_kwargs=None):
        super(ConvDropoutNormNonlin, self).__init__()
        self.nonlin_kwargs = nonlin_kwargs if nonlin_kwargs is not None else {'negative_slope': 1e-2, 'inplace': True}
        self.dropout_op_kwargs = dropout_op_kwargs if dropout_op_kwargs is not None else {'p': 0.5, 'inplace': True}
        self.norm_op_kwargs = norm_op_kwargs if norm_op_kwargs is not None else {'eps': 1e-5, 'affine': True,
                                                                                 'momentum': 0.1}
        self.conv_kwargs = conv_kwargs if conv_kwargs is not None else {'kernel_size': 3, 'stride': 1, 'padding': 1,
                                                                       'dilation': 1, 'bias': True}
        self.nonlin, self.dropout_op, self.conv_op, self.norm_op = nonlin, dropout_op, conv_op, norm_op
        self.conv = conv_op(input_channels, output_channels, **self.conv_kwargs)
        p = self.dropout_op_kwargs.get('p') if dropout_op is not None else None
        self.dropout = dropout_op(**self.dropout_op_kwargs) if (p is not None and p > 0) else None
        self.instnorm = norm_op(output_channels, **self.norm_op_kwargs)
        self.lrelu = nonlin(**self.nonlin_kwargs)

    def forward(self, x):
        x = self.conv(x)
        if self.dropout is not None:
            x = self.dropout(x)
        return self.lrelu(self.instnorm(x))


class ConvDropoutNonlinNorm(ConvDropoutNormNonlin):
    def forward(self, x):
        x = self.conv(x)
        if self.dropout is not None:
            x = self.dropout(x)
        return self.instnorm(self.lrelu(x))


class StackedConvLayers(nn.Module):
    """generic_UNet.py:81-144: `num_convs` blocks, `first_stride` on the first one only; children live in `.blocks`."""

    def __init__(self, input_feature_channels, output_feature_channels, num_convs, conv_op=nn.Conv2d, conv_kwargs=None,
                 norm_op=nn.BatchNorm2d, norm_op_kwargs=None, dropout_op=nn.Dropout2d, dropout_op_kwargs=None,
                 nonlin=nn.LeakyReLU, nonlin_kwargs=None, first_stride=None, basic_block=ConvDropoutNormNonlin):
        super(StackedConvLayers, self).__init__()
        self.input_channels, self.output_channels = input_feature_channels, output_feature_channels
        if conv_kwargs is None:
            conv_kwargs = {'kernel_size': 3, 'stride': 1, 'padding': 1, 'dilation': 1, 'bias': True}
        first_kwargs = conv_kwargs
        if first_stride is not None:
            first_kwargs = deepcopy(conv_kwargs)
            first_kwargs['stride'] = first_stride
        common = (norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs)
        layers = [basic_block(input_feature_channels, output_feature_channels, conv_op, first_kwargs, *common)]
        layers += [basic_block(output_feature_channels, output_feature_channels, conv_op, conv_kwargs, *common)
                   for _ in range(num_convs - 1)]
        self.blocks = nn.Sequential(*layers)

    def forward(self, x):
        return self.blocks(x)


def _is_native_block(blk):
    """True if a ConvDropoutNormNonlin leaf is something the kernels implement exactly."""
    c, n, a = blk.conv, blk.instnorm, blk.lrelu
    return (type(blk) is ConvDropoutNormNonlin and isinstance(c, nn.Conv3d) and blk.dropout is None
            and isinstance(n, nn.InstanceNorm3d) and n.affine and not n.track_running_stats and abs(n.eps - 1e-5) < 1e-12
            and isinstance(a, nn.LeakyReLU) and abs(a.negative_slope - 1e-2) < 1e-12
            and tuple(c.dilation) == (1, 1, 1) and c.groups == 1 and c.padding_mode == 'zeros'
            and all(k in (1, 3) for k in c.kernel_size)
            and tuple(c.padding) == tuple((k - 1) // 2 for k in c.kernel_size)
            and all(s in (1, 2) for s in c.stride))


class _UNetFunction(torch.autograd.Function):
    """The whole network as ONE autograd node: forward runs the kernel sequence and keeps a tape; backward replays it."""

    @staticmethod
    def forward(ctx, net, x, n_out, *params):
        tape = Tape() if any(p.requires_grad for p in params) else None
        net._engine.begin_step()
        feats = net._native_forward(x, tape)
        feats = feats[:n_out]
        ctx.net, ctx.tape, ctx.feats, ctx.params = net, tape, feats, params
        outs = tuple(f.as_ncdhw() for f in feats)
        ctx.mark_non_differentiable(*[o for o in outs if tape is None])
        return outs

    @staticmethod
    def backward(ctx, *grads):
        net, tape = ctx.net, ctx.tape
        eng = net._engine
        for f, g in zip(ctx.feats, grads):
            # a head whose d(logits) stayed with the loss (engine.lazy_heads) is seeded by the fused kernel itself
            if g is not None and f.buf.data_ptr() not in eng.lazy_heads:
                eng.seed_grad(tape, f, g)
        eng.run_backward(tape)
        out = []
        for p in ctx.params:
            g = tape.param_grads.get(id(p))
            if id(p) in tape.direct_done:
                # the kernels accumulated into the parameter's arena gradient slot: nothing for autograd to add
                assert g is None
                out.append(None)
                continue
            if g is None and p.requires_grad:
                g = torch.zeros_like(p)
            out.append(g if p.requires_grad else None)
        ctx.tape = None
        return (None, None, None) + tuple(out)


class Generic_UNet(SegmentationNetwork):
    DEFAULT_BATCH_SIZE_3D = 2
    DEFAULT_PATCH_SIZE_3D = (64, 192, 160)
    SPACING_FACTOR_BETWEEN_STAGES = 2
    BASE_NUM_FEATURES_3D = 30
    MAX_NUMPOOL_3D = 999
    MAX_NUM_FILTERS_3D = 320
    DEFAULT_PATCH_SIZE_2D = (256, 256)
    BASE_NUM_FEATURES_2D = 30
    DEFAULT_BATCH_SIZE_2D = 50
    MAX_NUMPOOL_2D = 999
    MAX_FILTERS_2D = 480
    use_this_for_batch_size_computation_2D = 19739648
    use_this_for_batch_size_computation_3D = 520000000

    def __init__(self, input_channels, base_num_features, num_classes, num_pool, num_conv_per_stage=2,
                 feat_map_mul_on_downscale=2, conv_op=nn.Conv2d, norm_op=nn.BatchNorm2d, norm_op_kwargs=None,
                 dropout_op=nn.Dropout2d, dropout_op_kwargs=None, nonlin=nn.LeakyReLU, nonlin_kwargs=None,
                 deep_supervision=True, dropout_in_localization=False, final_nonlin=softmax_helper,
                 weightInitializer=InitWeights_He(1e-2), pool_op_kernel_sizes=None, conv_kernel_sizes=None,
                 upscale_logits=False, convolutional_pooling=False, convolutional_upsampling=False,
                 max_num_features=None, basic_block=ConvDropoutNormNonlin, seg_output_use_bias=False,
                 internal_conv_bias=True, native_dtype=torch.float32, native_impl=0):
        """Same positional signature as the reference (generic_UNet.py:173-182); `native_dtype` / `native_impl` are
        keyword-only extensions selecting the kernel arithmetic (fp32 parity mode, bf16/fp16 tensor-core mode)."""
        super(Generic_UNet, self).__init__()
        self.convolutional_upsampling = convolutional_upsampling
        self.convolutional_pooling = convolutional_pooling
        self.upscale_logits = upscale_logits
        nonlin_kwargs = nonlin_kwargs if nonlin_kwargs is not None else {'negative_slope': 1e-2, 'inplace': True}
        dropout_op_kwargs = dropout_op_kwargs if dropout_op_kwargs is not None else {'p': 0.5, 'inplace': True}
        norm_op_kwargs = norm_op_kwargs if norm_op_kwargs is not None else {'eps': 1e-5, 'affine': True, 'momentum': 0.1}
        self.conv_kwargs = {'stride': 1, 'dilation': 1, 'bias': internal_conv_bias}
        self.nonlin, self.nonlin_kwargs = nonlin, nonlin_kwargs
        self.dropout_op, self.dropout_op_kwargs = dropout_op, dropout_op_kwargs
        self.norm_op, self.norm_op_kwargs = norm_op, norm_op_kwargs
        self.weightInitializer = weightInitializer
        self.conv_op = conv_op
        self.num_classes = num_classes
        self.final_nonlin = final_nonlin
        self._deep_supervision = deep_supervision
        self.do_ds = deep_supervision
        self.upsample_align_corners = True

        if conv_op == nn.Conv2d:
            upsample_mode, pool_op, transpconv, nd = 'bilinear', nn.MaxPool2d, nn.ConvTranspose2d, 2
        elif conv_op == nn.Conv3d:
            upsample_mode, pool_op, transpconv, nd = 'trilinear', nn.MaxPool3d, nn.ConvTranspose3d, 3
        else:
            raise ValueError("unknown convolution dimensionality, conv op: %s" % str(conv_op))
        if pool_op_kernel_sizes is None:
            pool_op_kernel_sizes = [(2,) * nd] * num_pool
        if conv_kernel_sizes is None:
            conv_kernel_sizes = [(3,) * nd] * (num_pool + 1)
        self.input_shape_must_be_divisible_by = np.prod(pool_op_kernel_sizes, 0, dtype=np.int64)
        self.pool_op_kernel_sizes = pool_op_kernel_sizes
        self.conv_kernel_sizes = conv_kernel_sizes
        self.conv_pad_sizes = [[1 if i == 3 else 0 for i in k] for k in conv_kernel_sizes]
        if max_num_features is None:
            max_num_features = self.MAX_NUM_FILTERS_3D if conv_op == nn.Conv3d else self.MAX_FILTERS_2D
        self.max_num_features = max_num_features

        def stack(cin, cout, n, level, first_stride=None, drop_kwargs=None):
            kw = dict(self.conv_kwargs)
            kw['kernel_size'] = self.conv_kernel_sizes[level]
            kw['padding'] = self.conv_pad_sizes[level]
            return StackedConvLayers(cin, cout, n, self.conv_op, kw, self.norm_op, self.norm_op_kwargs, self.dropout_op,
                                     drop_kwargs if drop_kwargs is not None else self.dropout_op_kwargs, self.nonlin,
                                     self.nonlin_kwargs, first_stride, basic_block=basic_block)

        context, localization, td, tu, seg = [], [], [], [], []
        feats_out, feats_in = base_num_features, input_channels
        for d in range(num_pool):
            stride = pool_op_kernel_sizes[d - 1] if (d != 0 and convolutional_pooling) else None
            context.append(stack(feats_in, feats_out, num_conv_per_stage, d, stride))
            if not convolutional_pooling:
                td.append(pool_op(pool_op_kernel_sizes[d]))
            feats_in = feats_out
            feats_out = min(int(np.round(feats_out * feat_map_mul_on_downscale)), self.max_num_features)

        # bottleneck: two stacks so that the last conv can change the width when upsampling is not convolutional
        stride = pool_op_kernel_sizes[-1] if convolutional_pooling else None
        final_feats = feats_out if convolutional_upsampling else context[-1].output_channels
        context.append(nn.Sequential(stack(feats_in, feats_out, num_conv_per_stage - 1, num_pool, stride),
                                     stack(feats_out, final_feats, 1, num_pool)))

        loc_drop = dict(self.dropout_op_kwargs)
        if not dropout_in_localization:
            loc_drop['p'] = 0.0
        for u in range(num_pool):
            from_down = final_feats
            from_skip = context[-(2 + u)].output_channels
            if u != num_pool - 1 and not convolutional_upsampling:
                final_feats = context[-(3 + u)].output_channels
            else:
                final_feats = from_skip
            k_up = pool_op_kernel_sizes[-(u + 1)]
            if convolutional_upsampling:
                tu.append(transpconv(from_down, from_skip, k_up, k_up, bias=False))
            else:
                tu.append(Upsample(scale_factor=k_up, mode=upsample_mode, align_corners=self.upsample_align_corners))
            level = len(self.conv_kernel_sizes) - (u + 1)
            localization.append(nn.Sequential(stack(from_skip * 2, from_skip, num_conv_per_stage - 1, level,
                                                    drop_kwargs=loc_drop),
                                              stack(from_skip, final_feats, 1, level, drop_kwargs=loc_drop)))
        for ds in range(len(localization)):
            seg.append(conv_op(localization[ds][-1].output_channels, num_classes, 1, 1, 0, 1, 1, seg_output_use_bias))

        self.upscale_logits_ops = []
        cum_upsample = np.cumprod(np.vstack(pool_op_kernel_sizes), axis=0)[::-1]
        for usl in range(num_pool - 1):
            if self.upscale_logits:
                self.upscale_logits_ops.append(Upsample(scale_factor=tuple(int(i) for i in cum_upsample[usl + 1]),
                                                        mode=upsample_mode, align_corners=self.upsample_align_corners))
            else:
                self.upscale_logits_ops.append(lambda x: x)

        # registration order is part of the checkpoint format (generic_UNet.py:366-370)
        self.conv_blocks_localization = nn.ModuleList(localization)
        self.conv_blocks_context = nn.ModuleList(context)
        self.td = nn.ModuleList(td)
        self.tu = nn.ModuleList(tu)
        self.seg_outputs = nn.ModuleList(seg)
        if self.upscale_logits:
            self.upscale_logits_ops = nn.ModuleList(self.upscale_logits_ops)
        if self.weightInitializer is not None:
            self.apply(self.weightInitializer)

        self._engine = Engine(native_dtype, native_impl)
        self._ops = None
        self._native_ok = self._check_native()

    # ---- native path -------------------------------------------------------------------------------------------------
    def set_native_dtype(self, dtype, impl=None):
        self._engine = Engine(dtype, self._engine.impl if impl is None else impl)

    def native_dtype(self):
        return self._engine.dtype

    def native_input_channels_padded(self):
        return pad_channels(self.conv_blocks_context[0].blocks[0].conv.in_channels)

    def _all_blocks(self):
        for m in list(self.conv_blocks_context) + list(self.conv_blocks_localization):
            stacks = list(m) if isinstance(m, nn.Sequential) else [m]
            for s in stacks:
                for b in s.blocks:
                    yield b

    def _check_native(self):
        if self.conv_op != nn.Conv3d or not self.convolutional_pooling or not self.convolutional_upsampling:
            return False
        if self.upscale_logits or self.num_classes > 64:
            return False
        if not all(_is_native_block(b) for b in self._all_blocks()):
            return False
        for t in self.tu:
            if not isinstance(t, nn.ConvTranspose3d) or t.bias is not None or tuple(t.kernel_size) != tuple(t.stride):
                return False
        return True

    def _build_ops(self):
        """ConvOp wrappers around the module parameters, in execution order."""
        def block_ops(blocks, split=0):
            ops = []
            for i, b in enumerate(blocks):
                c = b.conv
                ops.append((ConvOp(c.weight, c.bias, c.kernel_size, c.stride, split=split if i == 0 else 0),
                            b.instnorm.weight, b.instnorm.bias))
            return ops
        enc = [block_ops(list(s.blocks)) for s in list(self.conv_blocks_context)[:-1]]
        bott = block_ops([b for st in self.conv_blocks_context[-1] for b in st.blocks])
        dec, tus, heads = [], [], []
        for u in range(len(self.tu)):
            t = self.tu[u]
            tus.append(ConvOp(t.weight, None, t.kernel_size, t.stride, transposed=True))
            blocks = [b for st in self.conv_blocks_localization[u] for b in st.blocks]
            dec.append(block_ops(blocks, split=t.out_channels))
            h = self.seg_outputs[u]
            heads.append(ConvOp(h.weight, h.bias, h.kernel_size, h.stride))
        self._ops = dict(enc=enc, bott=bott, dec=dec, tu=tus, head=heads)
        PackGroup(collect_ops(self._ops))

    def _native_forward(self, x, tape, only_full_res=False):
        """Kernel sequence of Generic_UNet.forward (generic_UNet.py:379-401).  `x` is an NCDHW tensor or an NDHWC Feat.
        Returns logits Feats highest resolution first."""
        L.lib()  # fail loudly if the CUDA library is missing
        eng = self._engine
        if self._ops is None or self._ops['enc'][0][0][0].weight is not self.conv_blocks_context[0].blocks[0].conv.weight:
            self._build_ops()
        ops = self._ops
        f = x if isinstance(x, Feat) else eng.input_feat(x, compact=eng.use_c1(ops['enc'][0][0][0]))
        dev = f.buf.device
        # Frozen-trunk fast path (fine-tuning with only the segmentation heads trainable, nnUNetTrainerV2_warmup.py:
        # 122-124): when no parameter below the heads requires a gradient, the trunk records nothing on the tape and the
        # backward pass is the heads' weight gradients alone -- no data gradient, no InstanceNorm backward.
        ttape = tape
        if tape is not None and not any(p.requires_grad for n, p in self.named_parameters()
                                        if not n.startswith("seg_outputs.")):
            ttape = None
        head_dgrad = ttape is not None
        skips = []
        first = True
        mat = eng.materialize_inputs
        for stage in ops['enc']:
            for i, (op, g, b) in enumerate(stage):
                out = None
                last = i == len(stage) - 1
                if last:
                    # the stage output is the skip: it goes straight into the second half of the concat buffer --
                    # raw (norm-on-load kernels) or, for the TMA-fed tensor-core kernels, as the materialised activation
                    od = op.out_dims(f.dims)
                    planar = eng.planar_concat_ok(op.Cout_p, od)
                    if planar:  # two compact halves [2, B, D, H, W, Cp]: 0 = up-sampled features, 1 = this skip
                        cat = torch.empty((2,) + tuple(od) + (op.Cout_p,), dtype=eng.dtype, device=dev)
                    else:
                        cat = eng.new_buf(od, 2 * op.Cout_p, dev)
                    if not mat:
                        out = Feat(cat, op.Cout_p, op.Cout, op.Cout_p)
                f = eng.conv_norm(ttape, op, g, b, f, out, need_input_grad=not first)
                f.single_consumer = not last  # the stage output also feeds the decoder (skip connection)
                if last and mat:
                    f.act = eng.materialize(f, out=(Feat(cat[1], 0, op.Cout, op.Cout_p, planar=cat, half=1) if planar
                                                    else Feat(cat, op.Cout_p, op.Cout, op.Cout_p)))
                first = False
            skips.append(f)
            # data parallel: in the backward pass everything recorded AFTER this point (deeper encoder stages,
            # bottleneck, whole decoder, heads: > 99 % of the gradient bytes) is finished when this closure runs -- the
            # trainer starts the gradient all-reduce there, under the backward of the two widest encoder stages
            if ttape is not None and eng.backward_mark is not None and len(skips) == 2 and len(ops['enc']) > 2:
                ttape.closures.append(eng.backward_mark)
        for (op, g, b) in ops['bott']:
            f = eng.conv_norm(ttape, op, g, b, f)
            f.single_consumer = True  # next bottleneck conv / the first transposed conv
        logits = []
        nu = len(ops['tu'])
        for u in range(nu):
            skip = skips[-(u + 1)]
            top = ops['tu'][u]
            assert top.Cout_p == skip.Cp
            base = skip.act.planar if mat else None
            if base is not None:  # planar halves: the transposed conv writes half 0, the level's first conv reads both
                eng.conv_plain(ttape, top, f, Feat(base[0], 0, top.Cout, top.Cout_p, planar=base, half=0))
                f = Feat(base.view((-1,) + tuple(base.shape[2:])), 0, top.Cout + skip.C, 2 * skip.Cp, planar=base)
                cat = None
            else:
                cat = skip.act.buf if mat else skip.buf
                assert cat.shape[4] == 2 * skip.Cp
                # the transposed conv writes the first half of the same buffer: torch.cat (generic_UNet.py:392) vanishes
                eng.conv_plain(ttape, top, f, Feat(cat, 0, top.Cout, top.Cout_p))
            if base is not None:
                pass
            elif mat:
                f = Feat(cat, 0, top.Cout + skip.C, 2 * skip.Cp)
            else:
                ident = torch.zeros((cat.shape[0], skip.Cp, 4), dtype=torch.float32, device=dev)
                ident[:, :, 0] = 1.0
                ident[:, :, 2] = 1.0
                f = Feat(cat, 0, top.Cout + skip.C, 2 * skip.Cp, xform=torch.cat((ident, skip.xform), dim=1))
            nd = len(ops['dec'][u])
            for i, (op, g, b) in enumerate(ops['dec'][u]):
                f = eng.conv_norm(ttape, op, g, b, f)
                # the level's last conv feeds its head AND the next transposed conv -- except at full resolution
                f.single_consumer = i < nd - 1 or u == nu - 1
            if only_full_res and u != nu - 1:
                continue
            logits.append(eng.conv_plain(tape, ops['head'][u], f, need_input_grad=head_dgrad, head=True))
        return logits[::-1]

    def forward(self, x):
        if self._native_ok and x.is_cuda:
            want_ds = self._deep_supervision and self.do_ds
            params = tuple(self.parameters())
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                n_out = len(self.tu) if want_ds else 1
                outs = _UNetFunction.apply(self, x, n_out, *params)
            else:
                self._engine.begin_step()
                feats = self._native_forward(x, None, only_full_res=not want_ds)
                outs = tuple(f.as_ncdhw() for f in feats)
            outs = tuple(self.final_nonlin(o) for o in outs)
            if want_ds:
                return tuple([outs[0]] + [i(j) for i, j in zip(list(self.upscale_logits_ops)[::-1], outs[1:])])
            return outs[0]
        if self._native_ok and not x.is_cuda:
            raise L.Mtb200Error("Generic_UNet (MultiTalent configuration) runs on the native CUDA path only; got a %s "
                                "tensor. There is no CPU fallback." % x.device)
        return self._forward_torch(x)

    def _forward_torch(self, x):
        """Configurations outside the native path: same dataflow through the torch leaf modules."""
        skips, seg_outputs = [], []
        for d in range(len(self.conv_blocks_context) - 1):
            x = self.conv_blocks_context[d](x)
            skips.append(x)
            if not self.convolutional_pooling:
                x = self.td[d](x)
        x = self.conv_blocks_context[-1](x)
        for u in range(len(self.tu)):
            x = torch.cat((self.tu[u](x), skips[-(u + 1)]), dim=1)
            x = self.conv_blocks_localization[u](x)
            seg_outputs.append(self.final_nonlin(self.seg_outputs[u](x)))
        if self._deep_supervision and self.do_ds:
            return tuple([seg_outputs[-1]] + [i(j) for i, j in zip(list(self.upscale_logits_ops)[::-1],
                                                                  seg_outputs[:-1][::-1])])
        return seg_outputs[-1]

    def native_logits(self, tile: Feat) -> Feat:
        self._engine.begin_step()
        return self._native_forward(tile, None, only_full_res=True)[0]

    @staticmethod
    def compute_approx_vram_consumption(patch_size, num_pool_per_axis, base_num_features, max_num_features,
                                        num_modalities, num_classes, pool_op_kernel_sizes, deep_supervision=False,
                                        conv_per_stage=2):
        """generic_UNet.py:403-442 (planner constant; host arithmetic only)."""
        npool = len(pool_op_kernel_sizes)
        size = np.array(patch_size, dtype=np.int64)
        vox = np.prod(size, dtype=np.int64)
        total = np.int64((conv_per_stage * 2 + 1) * vox * base_num_features + num_modalities * vox + num_classes * vox)
        feat = base_num_features
        for p in range(npool):
            size = size // np.array(pool_op_kernel_sizes[p], dtype=np.int64)
            feat = min(feat * 2, max_num_features)
            nblocks = (conv_per_stage * 2 + 1) if p < (npool - 1) else conv_per_stage
            total += nblocks * np.prod(size, dtype=np.int64) * feat
            if deep_supervision and p < (npool - 2):
                total += np.prod(size, dtype=np.int64) * num_classes
        return total
