"""ResNet-18 / ResNet-50 on the sm_100a layers (BASELINE.json configs 2 and 4).

Architecture and ``state_dict`` keys follow the standard ImageNet-style ResNet
(7x7/2 stem, 3x3/2 max-pool, four stages, global average pool, linear head), so
a stock PyTorch ResNet ``state_dict`` of the same depth loads unchanged -- the
"checkpoint layout stays compatible" requirement.  With ``num_classes=10`` the
ResNet-18 float state is 11,191,242 elements.

Execution differs from a stock model: activations are bf16 NHWC, every
convolution is im2col + tcgen05 GEMM, BatchNorm fuses the residual add and the
ReLU of the block, parameters live in the flat arena.  The user-model contract
(``name``, ``__hash__``, ``train(X, y, n_epoch=...)``) comes from
``FederatedModule`` (reference demo.py:15-49).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Type

import torch
from torch import nn

from ..ops import nn as bnn
from .base import FederatedModule


# ---------------------------------------------------------------------------------------------------------------
# Hand-scheduled training step (no autograd engine).  The layer kernels are the ones the autograd Functions in
# ``ops/nn.py`` call -- their ``forward`` / ``backward`` bodies are driven directly through ``bnn.Ctx`` -- but the
# schedule is ours: the gradient of a block input travels as TWO pieces (main-branch dgrad, residual-branch
# gradient) that the consuming BatchNorm-backward kernel sums while loading, so the eight element-wise adds autograd
# would launch per step disappear; the downsample branch runs as a parallel branch of the captured graph.
# ---------------------------------------------------------------------------------------------------------------
def _conv_bn_fwd(conv: "bnn.Conv2d", bn: "bnn.BatchNorm2d", x, residual=None, after_conv=None):
    cc, cb = bnn.Ctx(), bnn.Ctx()
    stats = conv._fusable_stats(x)
    z = bnn._ConvFn.forward(cc, x, conv.weight, conv._w_bf16(), conv.kernel_size, conv.kernel_size, conv.stride,
                            conv.padding, None, stats, conv.flags_cfg)
    if after_conv is not None:
        after_conv()
    y = bnn._BNFn.forward(cb, z, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                          bn.eps, bn.momentum, bn.relu, True, bn.workspace, None, stats is not None)
    return y, (cc, cb)


def _conv_bn_bwd(ctxs, dy_a, dy_b=None, needs_dx=True):
    """-> (gradient w.r.t. the conv input, gradient w.r.t. the residual input or None)"""
    cc, cb = ctxs
    cc.needs_dx = needs_dx
    out = bnn._BNFn.backward(cb, dy_a, dy_b)
    dz, dres = out[0], out[1]
    dx = bnn._ConvFn.backward(cc, dz)[0]
    return dx, dres


# The stem (conv1 -> bn1 -> relu -> maxpool): BatchNorm + ReLU + max-pool are ONE kernel forward and two backward (the
# normalised 16x16 activation is never written: the pooled output carries the ReLU mask, BatchNorm's backward sums run
# over the pooled gradient -- csrc/norm.cu, "ResNet stem").  BATON_STEM_FUSED=0 keeps the separate kernels.
_STEM_FUSED = __import__("os").environ.get("BATON_STEM_FUSED", "1") != "0"


def _stem_fwd(conv, bn, pool, x, after_conv=None):
    """-> ``(pooled, conv ctx, z, argmax, mean, rstd)`` or ``None`` when the fused kernels do not apply."""
    stats = conv._fusable_stats(x)
    if (stats is None or bn.workspace is None or not bn.relu or not bn.training or bn.num_features % 8
            or bnn._grad_target(bn.weight) is None or bnn._grad_target(bn.bias) is None):
        return None
    cc = bnn.Ctx()
    z = bnn._ConvFn.forward(cc, x, conv.weight, conv._w_bf16(), conv.kernel_size, conv.kernel_size, conv.stride,
                            conv.padding, None, stats, conv.flags_cfg)
    if after_conv is not None:
        after_conv()
    c = bn.num_features
    out = bnn.F.bn_relu_maxpool(z, bn.workspace[: 2 * c], bnn._unwrap(bn.weight), bnn._unwrap(bn.bias), bn.running_mean,
                                bn.running_var, bn.num_batches_tracked, bn.eps, bn.momentum, pool.k, pool.stride, pool.pad)
    if out is None:
        raise RuntimeError("bn_relu_maxpool rejected a shape _stem_fwd accepted")
    p, arg, mean, rstd = out
    return p, cc, z, arg, mean, rstd


def _stem_bwd(saved, bn, pool, dy_a, dy_b=None):
    p, cc, z, arg, mean, rstd = saved
    c = bn.num_features
    gamma = bnn._unwrap(bn.weight)
    tg, tb = bnn._grad_target(gamma), bnn._grad_target(bnn._unwrap(bn.bias))
    dz = bnn.F.bn_maxpool_bwd(z, p, arg, dy_a.contiguous(), dy_b.contiguous() if dy_b is not None else None, gamma, mean,
                              rstd, bn.workspace[2 * c:], tg, tb, pool.k, pool.stride, pool.pad)
    if dz is None:
        raise RuntimeError("bn_maxpool_bwd rejected a shape bn_relu_maxpool accepted")
    cc.needs_dx = False
    bnn._ConvFn.backward(cc, dz)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None):
        super().__init__()
        self.conv1 = bnn.Conv2d(inplanes, planes, 3, stride, 1)
        self.bn1 = bnn.BatchNorm2d(planes, relu=True)
        self.conv2 = bnn.Conv2d(planes, planes, 3, 1, 1)
        self.bn2 = bnn.BatchNorm2d(planes, relu=True)      # relu(bn2(.) + identity), fused
        self.downsample = downsample

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = self.bn1(self.conv1(x))
        return self.bn2(self.conv2(out), identity)

    units = (("conv1", "bn1"), ("conv2", "bn2"))


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None):
        super().__init__()
        self.conv1 = bnn.Conv2d(inplanes, planes, 1, 1, 0)
        self.bn1 = bnn.BatchNorm2d(planes, relu=True)
        self.conv2 = bnn.Conv2d(planes, planes, 3, stride, 1)
        self.bn2 = bnn.BatchNorm2d(planes, relu=True)
        self.conv3 = bnn.Conv2d(planes, planes * 4, 1, 1, 0)
        self.bn3 = bnn.BatchNorm2d(planes * 4, relu=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = self.bn1(self.conv1(x))
        out = self.bn2(self.conv2(out))
        return self.bn3(self.conv3(out), identity)

    units = (("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3"))


class _Downsample(nn.Sequential):
    """``0`` = 1x1 strided conv, ``1`` = BatchNorm (same indices as the stock model)."""

    def __init__(self, inplanes: int, outplanes: int, stride: int):
        super().__init__(bnn.Conv2d(inplanes, outplanes, 1, stride, 0), bnn.BatchNorm2d(outplanes, relu=False))


class ResNet(FederatedModule):
    loss_kind = "ce"
    default_lr = 0.05
    default_batch_size = 128

    def __init__(self, block: Type[nn.Module], layers: Sequence[int], num_classes: int = 1000,
                 in_channels: int = 3, name: Optional[str] = None):
        super().__init__()
        if name:
            self.name = name
        self.inplanes = 64
        self.conv1 = bnn.Conv2d(in_channels, 64, 7, 2, 3)
        self.bn1 = bnn.BatchNorm2d(64, relu=True)
        self.maxpool = bnn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make_layer(block, 64, layers[0], 1)
        self.layer2 = self._make_layer(block, 128, layers[1], 2)
        self.layer3 = self._make_layer(block, 256, layers[2], 2)
        self.layer4 = self._make_layer(block, 512, layers[3], 2)
        self.avgpool = bnn.GlobalAvgPool()
        self.fc = bnn.Linear(512 * block.expansion, num_classes, out_fp32=True)
        self.stats_workspace = None
        self.zeroes_own_workspace = True   # forward() clears the statistics workspace itself
        for m in self.modules():   # zero-init the last BN of each block (standard recipe; keeps early training stable)
            if isinstance(m, BasicBlock):
                nn.init.zeros_(m.bn2.weight)
            elif isinstance(m, Bottleneck):
                nn.init.zeros_(m.bn3.weight)

    def _make_layer(self, block, planes: int, blocks: int, stride: int) -> nn.Sequential:
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = _Downsample(self.inplanes, planes * block.expansion, stride)
        layers: List[nn.Module] = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def set_precision(self, dtype: str = "bf16") -> "ResNet":
        """``"fp8"``: every convolution (forward, dgrad and wgrad GEMMs) runs on the block-scaled
        MXFP8 tensor-core path; BatchNorm, the classifier head and the optimizer stay bf16/fp32."""
        assert dtype in ("bf16", "fp8")
        for m in self.modules():
            if isinstance(m, bnn.Conv2d):
                m.fp8 = dtype == "fp8"
        self.compute_dtype = dtype
        return self

    def forward(self, x):
        """``x``: NHWC ``[N, H, W, C]`` (bf16 on CUDA)."""
        if self.stats_workspace is not None and self.training:
            self.stats_workspace.zero_()      # ONE memset per step: forward and backward statistic sums of every BN
        x = self.maxpool(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(self.avgpool(x))

    # ------------------------------------------------------------------ hand-scheduled step
    def _block_fwd(self, blk, x, tape):
        ds_ctx = None
        identity = x
        if blk.downsample is not None:
            with bnn.BRANCH.fork(x):                      # parallel graph branch: 1x1 conv + BN of the shortcut
                identity, ds_ctx = _conv_bn_fwd(blk.downsample[0], blk.downsample[1], x)
        out = x
        ctxs = []
        n = len(blk.units)
        for i, (cn, bnn_) in enumerate(blk.units):
            last = i == n - 1
            if last:
                bnn.BRANCH.join()                         # the shortcut must have landed before the residual add
            out, c = _conv_bn_fwd(getattr(blk, cn), getattr(blk, bnn_), out, identity if last else None)
            ctxs.append(c)
        tape.append((ctxs, ds_ctx))
        return out

    def _block_bwd(self, entry, pieces):
        """``pieces``: 1-2 tensors whose sum is the gradient of the block output -> pieces of the block-input gradient"""
        ctxs, ds_ctx = entry
        d, dres = _conv_bn_bwd(ctxs[-1], pieces[0], pieces[1] if len(pieces) > 1 else None)
        dx_ds = None
        if ds_ctx is not None:
            with bnn.BRANCH.fork(dres):                   # shortcut backward in parallel with the main branch
                dx_ds, _ = _conv_bn_bwd(ds_ctx, dres)
        for c in reversed(ctxs[:-1]):
            d, _ = _conv_bn_bwd(c, d)
        if ds_ctx is not None:
            bnn.BRANCH.join()
            return [d, dx_ds]
        return [d, dres]

    # BATON_SGD_OVERLAP=1: parameters from this prefix on receive their gradients first, their optimizer slice runs beside
    # the rest of the backward pass.  "layer1." = everything but the stem (SGD beside the stem's backward + weight gradient,
    # which leave most of the GPU idle); "layer3." = the deep layers only (88 % of a ResNet-18; measured neutral)
    tail_split_prefix = __import__("os").environ.get("BATON_SGD_SPLIT", "layer1.")

    def explicit_step(self, x, target, loss_acc=None, hooks=None):
        """Forward + loss + backward of one batch with parameter gradients accumulated into the arena (the same
        contract as ``loss.backward()`` on ``forward``); returns the device ``[mean loss, #correct]`` pair.
        CUDA + arena-adopted training mode only.

        ``hooks`` (optional, from the trainer): ``hooks.tail_grads_ready()`` is called as soon as the gradients of the
        deep layers (``tail_split_prefix`` onwards: 88 % of a ResNet-18's parameters) are complete, so their optimizer
        step can run beside the rest of the backward pass; ``hooks.before_tail_forward()`` is called before the first
        forward use of those weights."""
        F = bnn.F
        pre = self.tail_split_prefix or "layer3."
        tail_layer = int(pre[5]) - 1 if pre.startswith("layer") and pre[5:6].isdigit() else 2
        if self.stats_workspace is not None:
            self.stats_workspace.zero_()
        after_first = getattr(hooks, "after_first_gemm", None) if hooks is not None else None
        stem = cp = None
        fused_stem = _stem_fwd(self.conv1, self.bn1, self.maxpool, x, after_first) if _STEM_FUSED else None
        if fused_stem is not None:
            h = fused_stem[0]
        else:
            h, stem = _conv_bn_fwd(self.conv1, self.bn1, x, after_conv=after_first)
            cp = bnn.Ctx()
            h = bnn._MaxPoolFn.forward(cp, h, self.maxpool.k, self.maxpool.stride, self.maxpool.pad)
        tape = []
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4)):
            if li == tail_layer and hooks is not None and getattr(hooks, "_tail_pending", False):
                hooks.before_tail_forward()
            for blk in layer:
                h = self._block_fwd(blk, h, tape)
        n_head_blocks = sum(len(l) for l in (self.layer1, self.layer2, self.layer3, self.layer4)[:tail_layer])
        ca = None
        if h.shape[1] == 1 and h.shape[2] == 1:
            feat = h.reshape(h.shape[0], h.shape[3])
        else:
            ca = bnn.Ctx()
            feat = bnn._AvgPoolFn.forward(ca, h)
        fc = self.fc
        head = None
        tw, tb = bnn._grad_target(fc.weight), bnn._grad_target(fc.bias)
        if fc.out_features <= 32 and tw is not None and (fc.bias is None or tb is not None) and fc.act == 0:
            # classifier head (linear + softmax cross-entropy, forward and backward) in ONE launch
            head = F.linear_xent_head(feat.contiguous(), bnn._shadow(fc, "weight", fc.weight), fc.bias, target, tw, tb,
                                      acc=loss_acc)
        if head is not None:
            stats, d, _ = head
        else:
            cl = bnn.Ctx()
            logits = bnn._LinearFn.forward(cl, feat, fc.weight, fc.bias, bnn._shadow(fc, "weight", fc.weight), fc.act,
                                           fc.out_fp32, None, None)
            cl.needs_dx = True
            stats, dlogits = F.softmax_xent(logits.contiguous(), target, want_grad=True, acc=loss_acc)
            d = bnn._LinearFn.backward(cl, dlogits)[0]
        d = d.reshape(h.shape) if ca is None else bnn._AvgPoolFn.backward(ca, d)
        pieces = [d]
        for bi in range(len(tape) - 1, -1, -1):
            pieces = self._block_bwd(tape[bi], pieces)
            if bi == n_head_blocks and hooks is not None and getattr(hooks, "_split_active", 0):
                hooks.tail_grads_ready()
        if fused_stem is not None:
            _stem_bwd(fused_stem, self.bn1, self.maxpool, pieces[0], pieces[1] if len(pieces) > 1 else None)
        else:
            d = bnn._MaxPoolFn.backward(cp, pieces[0], pieces[1] if len(pieces) > 1 else None)[0]
            _conv_bn_bwd(stem, d, needs_dx=False)
        bnn.WGRAD.join()
        return stats

    # ------------------------------------------------------------------
    def build_workspace(self, device) -> torch.Tensor:
        """One fp32 buffer holding the forward/backward statistic sums of every
        BatchNorm layer (4*C each), zeroed by ONE memset per training step."""
        bns = [m for m in self.modules() if isinstance(m, bnn.BatchNorm2d)]
        total = sum(4 * m.num_features for m in bns)
        ws = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        for m in bns:
            m.workspace = ws[off: off + 4 * m.num_features]
            off += 4 * m.num_features
        self.stats_workspace = ws
        # convolution -> BatchNorm pairs: the conv GEMM's epilogue accumulates the batch statistics straight
        # into the BatchNorm's workspace, so the separate statistics pass over the activation disappears
        for parent in self.modules():
            for conv_name, bn_name in (("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3")):
                conv, bn = getattr(parent, conv_name, None), getattr(parent, bn_name, None)
                if isinstance(conv, bnn.Conv2d) and isinstance(bn, bnn.BatchNorm2d):
                    conv.bn_ws = bn.workspace
            if isinstance(parent, _Downsample):
                parent[0].bn_ws = parent[1].workspace
        return ws


def resnet18(num_classes: int = 10, **kw) -> ResNet:
    kw.setdefault("name", "resnet18")
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes=num_classes, **kw)


def resnet50(num_classes: int = 1000, **kw) -> ResNet:
    kw.setdefault("name", "resnet50")
    return ResNet(Bottleneck, [3, 4, 6, 3], num_classes=num_classes, **kw)
