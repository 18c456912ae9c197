"""Layers built on the sm_100a kernels, with hand-written backward passes.

Every layer keeps an fp32 master parameter (a view into the flat parameter
arena once ``ParamArena`` has adopted the model) and consumes a bf16 shadow
copy (``weight_bf16``, a view into the bf16 arena refreshed by the fused SGD
kernel).  Backward kernels accumulate parameter gradients *directly* into
``param.grad`` (views of the flat gradient arena) and report ``None`` to
autograd, so there is no per-parameter accumulate kernel and the optimizer is a
single pass over the arena.

Layout: activations are bf16 NHWC / ``[rows, features]``.  Conv weights are
logically ``[Cout, Cin, KH, KW]`` (state_dict compatible with stock PyTorch)
stored channels_last, i.e. physically ``[Cout, KH, KW, Cin]`` -- exactly the
K-major B operand of the implicit GEMM.

On a CPU tensor every layer falls back to the equivalent ``torch.nn.functional``
call so the control plane / tests run on a GPU-less host.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn
from torch.nn import functional as TF

from . import functional as F
from ._ext import load

BF16 = torch.bfloat16


class _P:
    """Opaque holder that carries a Parameter through ``Function.apply`` WITHOUT making it a
    differentiable input.  In arena mode the backward kernels accumulate into ``param.grad`` themselves,
    so autograd must not see the leaf: its cached AccumulateGrad node remembers the stream it was
    created on, and merely scheduling it inside a CUDA-graph capture makes the engine record an event on
    that (uncaptured) stream -> cudaErrorStreamCaptureIsolation.  A fresh zero-size ``anchor`` leaf
    stands in so backward still runs for layers whose data input needs no gradient."""
    __slots__ = ("p",)

    def __init__(self, p):
        self.p = p


def _unwrap(v):
    return v.p if isinstance(v, _P) else v


def _wrap(param, x):
    """(holder-or-param, anchor-or-None) for one layer call."""
    return param if (param is None or _grad_target(param) is None) else _P(param)


def _anchor(x, *params):
    if any(p is not None and _grad_target(p) is not None for p in params) and torch.is_grad_enabled():
        return torch.empty(0, device=x.device, dtype=torch.float32, requires_grad=True)
    return None


class Ctx:
    """Stand-in for the autograd context: lets a hand-scheduled training step (``models/resnet.py``
    ``ResNet.explicit_step``) call the ``forward`` / ``backward`` bodies of the Functions below directly, in its own
    order and with its own fusions (two-piece gradients, parallel branches), without the autograd engine."""

    def __init__(self):
        self.saved_tensors = ()

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors

    def mark_non_differentiable(self, *a):
        pass


class _WgradOverlap:
    """Weight-gradient GEMMs only feed the optimizer, so they are forked onto a side stream and overlap
    the dgrad -> BatchNorm-backward chain of the layers below (inside a captured graph this becomes a
    parallel branch).  Operand tensors are kept alive until the join so the caching allocator cannot hand
    their memory to the main stream early; the join is queued as an autograd end-of-backward callback and
    is also called by the trainer before the optimizer step."""

    def __init__(self):
        import os
        self.enabled = os.environ.get("BATON_WGRAD_OVERLAP", "1") != "0"
        self.max_flops = float(os.environ.get("BATON_WGRAD_OVERLAP_MAX_GFLOP", "4")) * 1e9
        self.streams = {}
        self.keep = []
        self.pending = False
        self.queued = False

    def mark(self, ref):
        """Record "the operands are ready" on the current stream.  Calling this BEFORE the dgrad GEMM is enqueued and
        :meth:`run` (with the returned token) AFTER it puts the dgrad kernel -- the one on the critical path -- first
        in the captured graph's launch order while the weight-gradient branch still only depends on what precedes it."""
        if not self.enabled or not ref.is_cuda:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(ref.device))
        return ev

    def run(self, fn, *keep, after=None):
        if not self.enabled or not keep[0].is_cuda:
            return fn()
        # dY[M, N] x X[M, K]: a GEMM that fills the machine on its own gains nothing from a parallel branch and
        # would only fight the persistent (one CTA per SM) dgrad kernels for SMs
        if 2.0 * keep[0].shape[0] * keep[0].shape[-1] * keep[1].shape[-1] > self.max_flops:
            return fn()
        dev = keep[0].device
        side = self.streams.get(dev)
        if side is None:
            side = self.streams[dev] = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        if after is not None:
            side.wait_event(after)
        else:
            side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn()
        self.keep.append(keep)
        self.pending = True
        if not self.queued:
            self.queued = True
            try:
                torch.autograd.Variable._execution_engine.queue_callback(self.join)
            except Exception:      # not inside a backward pass
                self.queued = False

    def join(self):
        self.queued = False
        if not self.pending:
            return
        for dev, side in self.streams.items():
            torch.cuda.current_stream(dev).wait_stream(side)
        self.keep.clear()
        self.pending = False


WGRAD = _WgradOverlap()


class _Branch:
    """Fork / join of an independent sub-chain (the shortcut conv + BN of a ResNet block, forward and backward) onto a
    side stream, so that inside a captured step it becomes a parallel graph branch.  These kernels use a fraction of
    the SMs and are latency bound, so two chains side by side cost the time of the longer one."""

    def __init__(self):
        import os
        self.enabled = os.environ.get("BATON_BRANCH_OVERLAP", "1") != "0"
        self.streams = {}
        self.keep = []
        self.pending = False

    def fork(self, *keep):
        import contextlib
        if not self.enabled or not keep[0].is_cuda:
            return contextlib.nullcontext()
        dev = keep[0].device
        side = self.streams.get(dev)
        if side is None:
            side = self.streams[dev] = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self.keep.append(keep)
        self.pending = True
        return torch.cuda.stream(side)

    def join(self):
        if not self.pending:
            return
        for dev, side in self.streams.items():
            torch.cuda.current_stream(dev).wait_stream(side)
        self.keep.clear()
        self.pending = False


BRANCH = _Branch()


def _grad_target(p: Optional[torch.Tensor]):
    """fp32 gradient buffer to accumulate into (arena view) or None."""
    if p is None or p.grad is None or p.grad.dtype != torch.float32:
        return None
    return p.grad


def _shadow(module: nn.Module, name: str, param: torch.Tensor, as2d: bool = True) -> torch.Tensor:
    """bf16 copy of a parameter.  Arena-adopted modules carry ``<name>_bf16``
    views that the SGD / FedAvg kernels keep in sync; otherwise cast on the fly."""
    sh = getattr(module, name + "_bf16", None)
    if sh is not None:
        return sh
    if param.dim() == 4:  # channels_last conv weight -> [Cout, KH*KW*Cin]
        src = param.detach().permute(0, 2, 3, 1).contiguous()
        return F.cast(src.view(param.shape[0], -1), BF16)
    return F.cast(param.detach().contiguous(), BF16)


# ================================================================================ Linear
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, w_bf16, act, out_fp32, flags_cfg, anchor):
        weight, bias = _unwrap(weight), _unwrap(bias)
        x2 = x.reshape(-1, x.shape[-1])
        kw = {}
        if flags_cfg is not None:
            kw = dict(flags=flags_cfg["flags"], flag_epoch=flags_cfg.get("epoch", 0), flag_elem_off=flags_cfg["elem_off"],
                      flag_tile_elems=flags_cfg["tile_elems"], flag_bias_off=flags_cfg.get("bias_off", -1),
                      force_bn=128, flag_epoch_word=flags_cfg.get("epoch_word"))
        y = F.gemm(x2, w_bf16, bias=bias, act=act if act != 2 else 0,
                   out_dtype=torch.float32 if out_fp32 else BF16, **kw)
        ctx.act = act
        pre = None
        if act == 2:  # GELU needs the pre-activation for backward
            pre = y
            y = F.gelu(pre)
        ctx.save_for_backward(x2, w_bf16, y if act == 1 else pre)
        ctx.weight, ctx.bias = weight, bias
        ctx.x_shape = x.shape
        ctx.needs_dx = x.requires_grad
        return y.view(*x.shape[:-1], w_bf16.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w_bf16, aux = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if dy2.dtype != BF16:
            dy2 = F.cast(dy2.contiguous(), BF16)
        elif not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if ctx.act == 1:
            dy2 = F.relu_bwd(aux, dy2)
        elif ctx.act == 2:
            dy2 = F.gelu_bwd(aux, dy2)
        weight, bias = ctx.weight, ctx.bias
        gw = gb = None
        # wgrad: dW[N, K] += dY^T[N, M] X[M, K]   (both operands MN-major, no transposes)
        tgt = _grad_target(weight)
        if tgt is not None:
            out2d = tgt.view(weight.shape[0], -1)
            WGRAD.run(lambda: F.gemm(dy2, x2, a_mn=True, b_mn=True, out=out2d, accumulate=True), dy2, x2)
        else:
            gw = F.gemm(dy2, x2, a_mn=True, b_mn=True, out_dtype=torch.float32, accumulate=True).view_as(weight)
        if bias is not None:
            tb = _grad_target(bias)
            if tb is not None:
                F.colsum_(dy2, tb, accumulate=True)
            else:
                gb = F.colsum_(dy2, torch.zeros_like(bias, dtype=torch.float32), accumulate=True)
        dx = None
        if ctx.needs_dx:
            # dgrad: dX[M, K] = dY[M, N] W[N, K]   (B = W is MN-major for this product)
            dx = F.gemm(dy2, w_bf16, b_mn=True).view(ctx.x_shape)
        return dx, gw, gb, None, None, None, None, None


class Linear(nn.Module):
    """``y = act(x W^T + b)``; ``act`` in {None, 'relu', 'gelu'} is fused into the GEMM epilogue."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, act: Optional[str] = None,
                 out_fp32: bool = False):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.act = {None: 0, "relu": 1, "gelu": 2}[act]
        self.out_fp32 = out_fp32
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        self.flags_cfg = None  # set by FedAvgSession for the first layer (bcast_gemm)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        if not x.is_cuda:
            y = TF.linear(x, self.weight.to(x.dtype), None if self.bias is None else self.bias.to(x.dtype))
            return TF.relu(y) if self.act == 1 else (TF.gelu(y, approximate="tanh") if self.act == 2 else y)
        if x.dtype != BF16:
            x = F.cast(x.contiguous(), BF16)
        if getattr(self, "fp8", False) and self.act == 0 and self.bias is None and not self.out_fp32:
            y = matmul_fp8(x.reshape(-1, x.shape[-1]), self, _shadow(self, "weight", self.weight), self.in_features)
            return y.view(*x.shape[:-1], self.out_features)
        cfg = self.flags_cfg
        if cfg is not None and cfg.get("epoch_word") is None:
            self.flags_cfg = None  # launch-constant epoch: one-shot, only the first GEMM after a round is gated
        return _LinearFn.apply(x, _wrap(self.weight, x), _wrap(self.bias, x), _shadow(self, "weight", self.weight),
                               self.act, self.out_fp32, cfg, _anchor(x, self.weight))


# ================================================================================ Conv2d (NHWC, implicit GEMM)
# implicit-GEMM convolution (TMA im2col operands; validated on B200 in round 2, profiles/r2_validate_experimental.txt).
# BATON_CONV_IGEMM=0 falls back to explicit im2col / col2im + GEMM.
_CONV_IGEMM = __import__("os").environ.get("BATON_CONV_IGEMM", "1") == "1"
_CONV_IGEMM_DGRAD = __import__("os").environ.get("BATON_CONV_IGEMM_DGRAD", "1") == "1"


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, w_bf16, kh, kw, stride, pad, anchor, stats=None, gate=None):
        """``gate``: bcast_gemm arrival-flag configuration of this layer's weights (first conv of the model only)."""
        weight = _unwrap(weight)
        n, h, w, c = x.shape
        cout = w_bf16.shape[0]
        # a k x k convolution (odd k, "same" padding) over a 1x1 feature map only ever sees its centre
        # tap -- every other tap multiplies zero padding.  Exact, and it turns the deepest ResNet stage
        # (32x32 inputs: layer4 is 1x1) into plain [N, Cin] x [Cout, Cin] GEMMs on a strided weight view:
        # no im2col / col2im, 9x less K.
        ctx.center = (h == 1 and w == 1 and kh == kw and kh % 2 == 1 and pad == kh // 2 and c % 8 == 0 and kh > 1)
        if ctx.center:
            col, ho, wo, kp = x.view(n, c), 1, 1, c
            wc = w_bf16.view(cout, kh * kw, c)[:, (kh // 2) * kw + kw // 2, :]
            y = F.gemm(col, wc, col_stats=stats)
        else:
            y = None
            ctx.igemm = False
            if kh == 1 and kw == 1 and stride == 1 and pad == 0 and c % 8 == 0:
                col, ho, wo, kp = x.view(n * h * w, c), h, w, c
            elif _CONV_IGEMM and c % 64 == 0:
                # A operand gathered by TMA im2col inside the GEMM, no col buffer (backward: implicit
                # wgrad from x itself)
                ho, wo, kp = F.conv_out_size(h, kh, stride, pad), F.conv_out_size(w, kw, stride, pad), kh * kw * c
                y = F.conv_igemm_fwd(x, w_bf16, kh, kw, stride, pad, col_stats=stats)
                col, ctx.igemm = x, y is not None
            if y is None:
                if not (kh == 1 and kw == 1 and stride == 1 and pad == 0 and c % 8 == 0):
                    col, ho, wo, kp = F.im2col(x, kh, kw, stride, pad)
                gk = {}
                if gate is not None:     # the TMA producer acquires the arrival flags over this layer's weights
                    gk = dict(flags=gate["flags"], flag_epoch_word=gate["epoch_word"], flag_elem_off=gate["elem_off"],
                              flag_tile_elems=gate["tile_elems"], force_bn=F.pick_bn(col.shape[0], w_bf16.shape[0]))
                y = F.gemm(col, w_bf16, col_stats=stats, **gk)
        ctx.save_for_backward(col, w_bf16)
        ctx.weight = weight
        ctx.geom = (n, h, w, c, kh, kw, stride, pad, ho, wo, kp)
        ctx.needs_dx = x.requires_grad
        return y.view(n, ho, wo, cout)

    @staticmethod
    def backward(ctx, dy):
        col, w_bf16 = ctx.saved_tensors
        n, h, w, c, kh, kw, stride, pad, ho, wo, kp = ctx.geom
        cout = w_bf16.shape[0]
        dy2 = dy.reshape(n * ho * wo, cout)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        weight = ctx.weight
        k_true = kh * kw * c
        gw = None
        tgt = _grad_target(weight)
        if ctx.center:
            tap = (kh // 2) * kw + kw // 2
            if tgt is None:
                gw = torch.zeros((cout, kh, kw, c), dtype=torch.float32, device=dy.device)
                g2d = gw.view(cout, kh * kw, c)[:, tap, :]
                gw = gw.permute(0, 3, 1, 2)
            else:
                full = tgt.permute(0, 2, 3, 1).reshape(cout, kh * kw, c)
                assert full.data_ptr() == tgt.data_ptr(), "conv weight grad must be channels_last in the arena"
                g2d = full[:, tap, :]
            if tgt is None:
                F.gemm(dy2, col, a_mn=True, b_mn=True, out=g2d, accumulate=True)
            else:
                WGRAD.run(lambda: F.gemm(dy2, col, a_mn=True, b_mn=True, out=g2d, accumulate=True), dy2, col)
            dx = None
            if ctx.needs_dx:
                wc = w_bf16.view(cout, kh * kw, c)[:, tap, :]
                dx = F.gemm(dy2, wc, b_mn=True).view(n, 1, 1, c)
            return dx, gw, None, None, None, None, None, None, None, None
        igemm = getattr(ctx, "igemm", False)          # col IS x: the weight gradient gathers im2col(x) on the fly
        tok = WGRAD.mark(dy2)      # the weight-gradient branch depends on what is enqueued so far, not on the dgrad below
        dx = None
        if ctx.needs_dx:
            if (_CONV_IGEMM and _CONV_IGEMM_DGRAD and stride == 1 and kh == kw and kh > 1 and kp == k_true
                    and w_bf16.shape[1] == k_true):
                # implicit dgrad: flipped-filter convolution of dy, no dcol buffer / col2im
                dx = F.conv_igemm_dgrad(dy2.view(n, ho, wo, cout), w_bf16, (n, h, w, c), kh, kw, pad)
            if dx is None:
                dcol = F.gemm(dy2, w_bf16, b_mn=True)  # [M, kp]
                if kh == 1 and kw == 1 and stride == 1 and pad == 0 and c % 8 == 0:
                    dx = dcol.view(n, h, w, c)
                else:
                    dx = F.col2im(dcol, (n, h, w, c), kh, kw, stride, pad, ho, wo)
        if tgt is not None:
            # the arena view is channels_last: physical [Cout, KH, KW, Cin] == [Cout, K]
            out2d = tgt.permute(0, 2, 3, 1).reshape(cout, k_true) if tgt.dim() == 4 else tgt.view(cout, k_true)
            assert out2d.data_ptr() == tgt.data_ptr(), "conv weight grad must be channels_last in the arena"
            if igemm:
                WGRAD.run(lambda: F.conv_igemm_wgrad_(dy2, col, out2d, kh, kw, stride, pad), dy2, col, after=tok)
            else:
                WGRAD.run(lambda: F.gemm(dy2, col, a_mn=True, b_mn=True, out=out2d, accumulate=True, n_valid=k_true),
                          dy2, col, after=tok)
        elif igemm:
            g2 = torch.zeros((cout, k_true), dtype=torch.float32, device=dy.device)
            F.conv_igemm_wgrad_(dy2, col, g2, kh, kw, stride, pad)
            gw = g2.view(cout, kh, kw, c).permute(0, 3, 1, 2)
        else:
            g2 = F.gemm(dy2, col, a_mn=True, b_mn=True, out_dtype=torch.float32, accumulate=True, n_valid=k_true)
            gw = g2.view(cout, kh, kw, c).permute(0, 3, 1, 2)
        return dx, gw, None, None, None, None, None, None, None, None


class Conv2d(nn.Module):
    """NHWC convolution (no bias -- every conv in ResNet is followed by BatchNorm)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int = 1, padding: int = 0):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        self.weight = nn.Parameter(w.contiguous(memory_format=torch.channels_last))
        self.k_true = kernel_size * kernel_size * in_channels
        self.kp = F.round_up(self.k_true, 8)
        self.flags_cfg = None   # bcast_gemm: set by FedAvgSession.gate_first_conv for the first conv of the model

    def _w_bf16(self):
        sh = getattr(self, "weight_bf16", None)
        if sh is None:
            sh = _shadow(self, "weight", self.weight)
        sh = sh.view(self.out_channels, self.k_true)
        if self.kp != self.k_true:  # K not a multiple of 8 (7x7x3 stem): zero-padded copy for TMA
            sh = F.pad_rows(sh, self.kp, gate=self.flags_cfg)
        return sh

    def forward(self, x):
        if not x.is_cuda:
            y = TF.conv2d(x.permute(0, 3, 1, 2), self.weight.to(x.dtype), None, self.stride, self.padding)
            return y.permute(0, 2, 3, 1)
        if getattr(self, "fp8", False):
            return conv2d_fp8(x, self)
        stats = self._fusable_stats(x)
        y = _ConvFn.apply(x, _wrap(self.weight, x), self._w_bf16(), self.kernel_size, self.kernel_size,
                          self.stride, self.padding, _anchor(x, self.weight), stats, self.flags_cfg)
        if stats is not None:
            y._bn_stats_ws = stats      # tells the BatchNorm that owns this workspace to skip its statistics pass
        return y

    def _fusable_stats(self, x):
        """The following BatchNorm's statistics workspace (``bn_ws``, linked by the model) if this call's GEMM
        can accumulate the batch statistics in its epilogue; ``None`` otherwise."""
        ws = getattr(self, "bn_ws", None)
        if ws is None or not self.training or not torch.is_grad_enabled():
            return None
        n, h, w, c = x.shape
        k, s_, p_ = self.kernel_size, self.stride, self.padding
        ho, wo = (h + 2 * p_ - k) // s_ + 1, (w + 2 * p_ - k) // s_ + 1
        centre = h == 1 and w == 1 and k % 2 == 1 and p_ == k // 2 and c % 8 == 0 and k > 1
        kdim = c if (centre or (k == 1 and c % 8 == 0)) else F.round_up(k * k * c, 8)
        cout = self.weight.shape[0]
        return ws[: 2 * cout] if F.gemm_stats_fusable(n * ho * wo, cout, kdim) else None


# ================================================================================ BatchNorm (+residual +ReLU)
_BN_BWD_FUSED = __import__("os").environ.get("BATON_BN_BWD_FUSED", "0") == "1"   # grid-barrier variant: validated, slower
# single-kernel BatchNorm backward, one thread-block cluster per 16-channel slice (csrc/norm.cu)
_BN_BWD_CLUSTER = __import__("os").environ.get("BATON_BN_BWD_CLUSTER", "1") == "1"
_BN_BWD_MAX_CLUSTER = int(__import__("os").environ.get("BATON_BN_BWD_MAX_CLUSTER", "16"))
_GRID_BARRIERS = {}


def _grid_barrier_words(device):
    """{count, generation} words of the device-wide barrier used by the single-kernel BatchNorm backward."""
    buf = _GRID_BARRIERS.get(device)
    if buf is None:
        buf = _GRID_BARRIERS[device] = torch.zeros(2, dtype=torch.int32, device=device)
    return buf


class _BNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, gamma, beta, rmean, rvar, nbt, eps, momentum, relu, training, ws, anchor,
                stats_ready=False):
        gamma, beta = _unwrap(gamma), _unwrap(beta)
        C_ = load()
        c = x.shape[-1]
        rows = x.numel() // c
        y = torch.empty_like(x)
        if ws is None:
            ws = torch.zeros(4 * c, dtype=torch.float32, device=x.device)
        sums_f, sums_b = ws[: 2 * c], ws[2 * c:]
        save_mean = torch.empty(c, dtype=torch.float32, device=x.device)
        save_rstd = torch.empty(c, dtype=torch.float32, device=x.device)
        if training and not stats_ready:     # stats_ready: the producing GEMM's epilogue already accumulated them
            C_.bn_stats(x, sums_f, rows, c)
        C_.bn_apply(x, residual, y, sums_f, gamma, beta, rmean, rvar, save_mean, save_rstd, nbt, rows, c, eps,
                    momentum, relu, training)
        ctx.save_for_backward(x, y, save_mean, save_rstd)
        ctx.gamma, ctx.beta, ctx.sums_b = gamma, beta, sums_b
        ctx.relu, ctx.has_res, ctx.rows, ctx.c = relu, residual is not None, rows, c
        return y

    @staticmethod
    def backward(ctx, dy, dy_b=None):
        """``dy_b``: optional second piece of the incoming gradient (``dy + dy_b``), summed inside the kernel --
        only the hand-scheduled step passes it (autograd sums gradients itself)."""
        C_ = load()
        x, y, mean, rstd = ctx.saved_tensors
        if not dy.is_contiguous():
            dy = dy.contiguous()
        gamma, beta = ctx.gamma, ctx.beta
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if ctx.has_res else None
        tg, tb = _grad_target(gamma), _grad_target(beta)
        gg = gb = None
        if tg is None:
            gg = tg = torch.zeros(ctx.c, dtype=torch.float32, device=x.device)
        if tb is None:
            gb = tb = torch.zeros(ctx.c, dtype=torch.float32, device=x.device)
        done = False
        if _BN_BWD_CLUSTER:
            done = C_.bn_bwd_cluster(x, y, dy, dy_b, dx, dres, gamma, mean, rstd, tg, tb, ctx.rows, ctx.c, ctx.relu,
                                     _BN_BWD_MAX_CLUSTER)
        if not done and dy_b is not None:
            dy = F.add(dy, dy_b.contiguous())
        if not done and _BN_BWD_FUSED:       # reduce + device-wide barrier + apply in one kernel
            done = C_.bn_bwd_fused(x, y, dy, dx, dres, gamma, mean, rstd, ctx.sums_b, tg, tb, ctx.rows, ctx.c, ctx.relu,
                                   _grid_barrier_words(x.device))
        if not done:
            C_.bn_bwd_reduce(x, y, dy, mean, rstd, ctx.sums_b, ctx.rows, ctx.c, ctx.relu)
            C_.bn_bwd_apply(x, y, dy, dx, dres, gamma, mean, rstd, ctx.sums_b, tg, tb, ctx.rows, ctx.c, ctx.relu)
        if ctx.has_res and dres is None:
            dres = dy
        return dx, dres, gg, gb, None, None, None, None, None, None, None, None, None, None


class BatchNorm2d(nn.Module):
    """BatchNorm over the channel (last) axis of an NHWC tensor with the residual add
    and ReLU of a ResNet block fused into the same pass:  ``relu(bn(x) + residual)``."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, relu: bool = False):
        super().__init__()
        self.num_features, self.eps, self.momentum, self.relu = num_features, eps, momentum, relu
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.workspace = None  # [4*C] fp32 slice of the model-wide stats workspace (zeroed once per step)

    def forward(self, x, residual=None):
        if not x.is_cuda:
            xn = x.permute(0, 3, 1, 2)
            y = TF.batch_norm(xn, self.running_mean, self.running_var, self.weight, self.bias, self.training,
                              self.momentum, self.eps).permute(0, 2, 3, 1)
            if self.training:
                self.num_batches_tracked += 1
            if residual is not None:
                y = y + residual
            return TF.relu(y) if self.relu else y
        ws = self.workspace
        if ws is None:
            ws = torch.zeros(4 * self.num_features, dtype=torch.float32, device=x.device)
        fused = getattr(x, "_bn_stats_ws", None)      # set by the producing Conv2d when its GEMM took the statistics
        ready = (fused is not None and self.workspace is not None and self.training
                 and fused.data_ptr() == self.workspace.data_ptr())
        return _BNFn.apply(x, residual, _wrap(self.weight, x), _wrap(self.bias, x), self.running_mean,
                           self.running_var, self.num_batches_tracked if self.training else None, self.eps,
                           self.momentum, self.relu, self.training, ws, _anchor(x, self.weight), ready)


# ================================================================================ pooling / misc
class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        y, arg = F.maxpool(x, k, stride, pad)
        ctx.save_for_backward(arg)
        ctx.cfg = (tuple(x.shape), k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy, dy_b=None):
        (arg,) = ctx.saved_tensors
        shape, k, stride, pad = ctx.cfg
        return F.maxpool_bwd(dy.contiguous(), arg, shape, k, stride, pad, dy_b=dy_b), None, None, None


class MaxPool2d(nn.Module):
    def __init__(self, kernel_size: int = 3, stride: int = 2, padding: int = 1):
        super().__init__()
        self.k, self.stride, self.pad = kernel_size, stride, padding

    def forward(self, x):
        if not x.is_cuda:
            return TF.max_pool2d(x.permute(0, 3, 1, 2), self.k, self.stride, self.pad).permute(0, 2, 3, 1)
        return _MaxPoolFn.apply(x, self.k, self.stride, self.pad)


class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return F.avgpool(x)

    @staticmethod
    def backward(ctx, dy):
        return F.avgpool_bwd(dy.contiguous(), ctx.shape)


class GlobalAvgPool(nn.Module):
    """NHWC ``[N,H,W,C] -> [N,C]`` (a view when H = W = 1, the 32x32-input ResNet case)."""

    def forward(self, x):
        if x.shape[1] == 1 and x.shape[2] == 1:
            return x.reshape(x.shape[0], x.shape[3])
        if not x.is_cuda:
            return x.mean(dim=(1, 2))
        return _AvgPoolFn.apply(x)


# ================================================================================ LayerNorm / GELU / softmax
class _LNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, gamma, beta, eps, anchor):
        gamma, beta = _unwrap(gamma), _unwrap(beta)
        C_ = load()
        c = x.shape[-1]
        rows = x.numel() // c
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        C_.layernorm_fwd(x, residual, y, gamma, beta, mean, rstd, rows, c, eps)
        pre = x if residual is None else F.add(x, residual)
        ctx.save_for_backward(pre, mean, rstd)
        ctx.gamma, ctx.beta, ctx.has_res, ctx.rows, ctx.c = gamma, beta, residual is not None, rows, c
        return y

    @staticmethod
    def backward(ctx, dy):
        C_ = load()
        pre, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(pre)
        tg, tb = _grad_target(ctx.gamma), _grad_target(ctx.beta)
        gg = gb = None
        if tg is None:
            gg = tg = torch.zeros(ctx.c, dtype=torch.float32, device=dy.device)
        if tb is None:
            gb = tb = torch.zeros(ctx.c, dtype=torch.float32, device=dy.device)
        C_.layernorm_bwd(pre, dy, dx, ctx.gamma, mean, rstd, tg, tb, ctx.rows, ctx.c)
        return dx, (dx if ctx.has_res else None), gg, gb, None, None


class LayerNorm(nn.Module):
    """``LN(x + residual)`` over the last axis (residual optional)."""

    def __init__(self, normalized_shape: int, eps: float = 1e-12):
        super().__init__()
        self.c, self.eps = normalized_shape, eps
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))

    def forward(self, x, residual=None):
        if not x.is_cuda:
            if residual is not None:
                x = x + residual
            return TF.layer_norm(x, (self.c,), self.weight.to(x.dtype), self.bias.to(x.dtype), self.eps)
        return _LNFn.apply(x.contiguous(), None if residual is None else residual.contiguous(),
                           _wrap(self.weight, x), _wrap(self.bias, x), self.eps, _anchor(x, self.weight))


class _SoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        c = x.shape[-1]
        y = torch.empty_like(x)
        load().softmax_fwd(x, y, x.numel() // c, c, scale)
        ctx.save_for_backward(y)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        c = y.shape[-1]
        dx = torch.empty_like(y)
        load().softmax_bwd(y, dy.contiguous(), dx, y.numel() // c, c, ctx.scale)
        return dx, None


def softmax(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """``softmax(scale * x)`` over the last axis."""
    if not x.is_cuda:
        return torch.softmax(x * scale, dim=-1)
    return _SoftmaxFn.apply(x.contiguous(), scale)


# ================================================================================ losses
class _XentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        acc, dl = F.softmax_xent(logits, target, want_grad=True)
        ctx.save_for_backward(dl)
        ctx.mark_non_differentiable(acc)
        return acc[0].clone(), acc

    @staticmethod
    def backward(ctx, g, _unused):
        (dl,) = ctx.saved_tensors
        return dl * g.to(dl.dtype), None


def cross_entropy(logits: torch.Tensor, target: torch.Tensor):
    """Fused softmax + NLL + gradient.  Returns ``(loss, stats)`` with
    ``stats = [mean loss, #correct]`` on the device."""
    if not logits.is_cuda:
        loss = TF.cross_entropy(logits.float(), target)
        hits = (logits.argmax(-1) == target).sum().float()
        return loss, torch.stack([loss.detach(), hits])
    return _XentFn.apply(logits.contiguous(), target)


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        acc, dp = F.mse(pred, target, want_grad=True)
        ctx.save_for_backward(dp)
        return acc[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dp,) = ctx.saved_tensors
        return dp * g.to(dp.dtype), None


def mse_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    if not pred.is_cuda:
        return TF.mse_loss(pred.float(), target.float().reshape(pred.shape))
    return _MseFn.apply(pred.contiguous(), target.reshape(pred.shape))


# ================================================================================ attention / embedding (BERT)
_FUSED_ATTN = __import__("os").environ.get("BATON_FUSED_ATTN", "1") == "1"   # validated on hardware; +4.5 % on BERT-base


class _AttnFn(torch.autograd.Function):
    """Multi-head self-attention core on a packed ``qkv [B*S, 3*H*dh]`` buffer: four strided-batched
    tcgen05 GEMMs + the row-softmax kernel forward, five GEMMs + softmax backward; Q/K/V and their
    gradients are addressed in place through 4-D TMA maps (no head split / merge copies)."""

    @staticmethod
    def forward(ctx, qkv, B, S, H, dh, mask_bias=None):
        D = H * dh
        if _FUSED_ATTN and S == 128 and dh == 64 and mask_bias is None:
            # single-kernel forward (csrc/attention.cu): scores stay in TMEM, P is written once
            probs = torch.empty((B * H * S, S), dtype=BF16, device=qkv.device)
            out = torch.empty((B * S, D), dtype=BF16, device=qkv.device)
            if load().attention_fwd(qkv, out, probs, B, S, H, dh, 1.0 / math.sqrt(dh)):
                ctx.save_for_backward(qkv, probs)
                ctx.dims = (B, S, H, dh)
                return out
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        scores = torch.empty((B * H * S, S), dtype=BF16, device=qkv.device)
        F.gemm_batched(q, k, scores, M=S, N=S, K=dh, lda=3 * D, ldb=3 * D, ldd=S, a_mn=False, b_mn=False,
                       n_outer=B, n_inner=H, a_strides=(S * 3 * D, dh), b_strides=(S * 3 * D, dh),
                       d_strides=(H * S * S, S * S), alpha=1.0 / math.sqrt(dh))
        if mask_bias is not None:        # additive key-padding mask [B, S] (0 / large negative), broadcast over heads and queries
            scores.view(B, H, S, S).add_(mask_bias.to(scores.dtype).view(B, 1, 1, S))
        probs = torch.empty_like(scores)
        load().softmax_fwd(scores, probs, B * H * S, S, 1.0)
        out = torch.empty((B * S, D), dtype=BF16, device=qkv.device)
        F.gemm_batched(probs, v, out, M=S, N=dh, K=S, lda=S, ldb=3 * D, ldd=D, a_mn=False, b_mn=True,
                       n_outer=B, n_inner=H, a_strides=(H * S * S, S * S), b_strides=(S * 3 * D, dh),
                       d_strides=(S * D, dh))
        ctx.save_for_backward(qkv, probs)
        ctx.dims = (B, S, H, dh)
        ctx.masked = mask_bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, probs = ctx.saved_tensors
        B, S, H, dh = ctx.dims
        D = H * dh
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        if _FUSED_ATTN and S == 128 and dh == 64 and not getattr(ctx, "masked", False):
            # single-kernel backward (csrc/attention.cu): dP / dS never leave the SM
            if load().attention_bwd(qkv, dout, probs, dqkv, B, S, H, dh, 1.0 / math.sqrt(dh)):
                return dqkv, None, None, None, None, None
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        dq, dk, dv = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
        bh = (H * S * S, S * S)
        pk = (S * 3 * D, dh)
        # dV = P^T dO
        F.gemm_batched(probs, dout, dv, M=S, N=dh, K=S, lda=S, ldb=D, ldd=3 * D, a_mn=True, b_mn=True,
                       n_outer=B, n_inner=H, a_strides=bh, b_strides=(S * D, dh), d_strides=pk)
        # dP = dO V^T
        dprobs = torch.empty_like(probs)
        F.gemm_batched(dout, v, dprobs, M=S, N=S, K=dh, lda=D, ldb=3 * D, ldd=S, a_mn=False, b_mn=False,
                       n_outer=B, n_inner=H, a_strides=(S * D, dh), b_strides=pk, d_strides=bh)
        dscores = torch.empty_like(probs)
        load().softmax_bwd(probs, dprobs, dscores, B * H * S, S, 1.0)
        alpha = 1.0 / math.sqrt(dh)
        # dQ = alpha dS K ; dK = alpha dS^T Q
        F.gemm_batched(dscores, k, dq, M=S, N=dh, K=S, lda=S, ldb=3 * D, ldd=3 * D, a_mn=False, b_mn=True,
                       n_outer=B, n_inner=H, a_strides=bh, b_strides=pk, d_strides=pk, alpha=alpha)
        F.gemm_batched(dscores, q, dk, M=S, N=dh, K=S, lda=S, ldb=3 * D, ldd=3 * D, a_mn=True, b_mn=True,
                       n_outer=B, n_inner=H, a_strides=bh, b_strides=pk, d_strides=pk, alpha=alpha)
        return dqkv, None, None, None, None, None


def attention(qkv: torch.Tensor, B: int, S: int, H: int, dh: int, mask_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``softmax(Q K^T / sqrt(dh) + mask_bias) V`` for packed ``qkv [B*S, 3*H*dh]`` -> ``[B*S, H*dh]``.
    ``mask_bias``: optional additive key mask ``[B, S]`` (0 = attend, large negative = padding)."""
    if not qkv.is_cuda:
        D = H * dh
        q, k, v = (t.reshape(B, S, H, dh).transpose(1, 2) for t in qkv.split(D, dim=-1))
        sc = q @ k.transpose(-1, -2) / math.sqrt(dh)
        if mask_bias is not None:
            sc = sc + mask_bias.to(sc.dtype).view(B, 1, 1, S)
        p = torch.softmax(sc, dim=-1)
        return (p @ v).transpose(1, 2).reshape(B * S, D)
    return _AttnFn.apply(qkv.contiguous(), B, S, H, dh, mask_bias)


class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, table_bf16, ids, anchor):
        table = _unwrap(table)
        ctx.save_for_backward(ids)
        ctx.table = table
        return F.gather_rows(table_bf16, ids)

    @staticmethod
    def backward(ctx, dy):
        (ids,) = ctx.saved_tensors
        table = ctx.table
        tgt = _grad_target(table)
        g = None
        if tgt is None:
            g = tgt = torch.zeros_like(table, dtype=torch.float32)
        F.embedding_bwd_(dy.contiguous(), ids, tgt)
        return g, None, None, None


class Embedding(nn.Module):
    """Lookup in the bf16 shadow of an fp32 table; gradient scattered with fp32 atomics."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(num_embeddings, embedding_dim) * 0.02)

    def forward(self, ids):
        flat = ids.reshape(-1)
        if not ids.is_cuda:
            return TF.embedding(flat, self.weight)
        return _EmbedFn.apply(_wrap(self.weight, flat), _shadow(self, "weight", self.weight), flat,
                              _anchor(flat, self.weight))


# ================================================================================ MXFP8 matmul (block-scaled fp8 training)
class _MatmulFp8Fn(torch.autograd.Function):
    """``y = x w^T`` with every GEMM of the layer -- forward, dgrad, wgrad -- on the block-scaled
    fp8 tensor-core path.  Each operand is quantised (e4m3 + UE8M0 scale per 32 elements) along the
    reduction dimension of the GEMM that consumes it; the transposed operands of dgrad / wgrad come
    out of the fused quantise+transpose kernel, so the GEMM kernel only ever sees K-major inputs.
    The weight gradient is accumulated in fp32 straight into the gradient arena."""

    @staticmethod
    def forward(ctx, x2, weight, w_bf16, k_true, anchor):
        weight = _unwrap(weight)
        K = x2.shape[1]
        xq, sx = F.quant_mx_rows(x2)
        wq, sw = F.quant_mx_rows(w_bf16)
        y = F.gemm_fp8(xq, sx, wq, sw, K)
        ctx.save_for_backward(x2, w_bf16)
        ctx.weight, ctx.k_true = weight, k_true
        ctx.needs_dx = x2.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, w_bf16 = ctx.saved_tensors
        dy = dy.contiguous()
        M, N = dy.shape
        K = x2.shape[1]
        weight, k_true = ctx.weight, ctx.k_true
        # wgrad: dW[N, K] = dY^T[N, M] X^T[K, M]^T   (reduction over M)
        dyt, sdyt = F.quant_mx_cols(dy)
        xt, sxt = F.quant_mx_cols(x2)
        gw = None
        tgt = _grad_target(weight)
        if tgt is not None:
            out2d = tgt.permute(0, 2, 3, 1).reshape(N, k_true) if tgt.dim() == 4 else tgt.view(N, k_true)
            F.gemm_fp8(dyt, sdyt, xt, sxt, M, out=out2d, accumulate=True, n_valid=k_true)
        else:
            g2 = F.gemm_fp8(dyt, sdyt, xt, sxt, M, out_dtype=torch.float32, accumulate=True, n_valid=k_true)
            gw = g2.view(weight.shape[0], *weight.shape[2:], weight.shape[1]).permute(0, 3, 1, 2) if weight.dim() == 4 \
                else g2.view_as(weight)
        dx = None
        if ctx.needs_dx:
            # dgrad: dX[M, K] = dY[M, N] (W^T)[K, N]^T   (reduction over N)
            dyq, sdy = F.quant_mx_rows(dy)
            wt, swt = F.quant_mx_cols(w_bf16)
            dx = F.gemm_fp8(dyq, sdy, wt, swt, N)
        return dx, gw, None, None, None


def matmul_fp8(x2: torch.Tensor, module: nn.Module, w_bf16: torch.Tensor, k_true: int) -> torch.Tensor:
    return _MatmulFp8Fn.apply(x2, _wrap(module.weight, x2), w_bf16, k_true, _anchor(x2, module.weight))


class _Im2colFn(torch.autograd.Function):
    """im2col as its own differentiable op (its adjoint is col2im) so the fp8 matmul above can sit
    between it and the BatchNorm that follows."""

    @staticmethod
    def forward(ctx, x, kh, kw, stride, pad):
        col, ho, wo, kp = F.im2col(x, kh, kw, stride, pad)
        ctx.geom = (tuple(x.shape), kh, kw, stride, pad, ho, wo)
        return col

    @staticmethod
    def backward(ctx, dcol):
        shape, kh, kw, stride, pad, ho, wo = ctx.geom
        return F.col2im(dcol.contiguous(), shape, kh, kw, stride, pad, ho, wo), None, None, None, None


def conv2d_fp8(x: torch.Tensor, conv: "Conv2d") -> torch.Tensor:
    """NHWC convolution with all three GEMMs in MXFP8."""
    n, h, w, c = x.shape
    k, s, p = conv.kernel_size, conv.stride, conv.padding
    ho, wo = F.conv_out_size(h, k, s, p), F.conv_out_size(w, k, s, p)
    if k == 1 and s == 1 and p == 0 and c % 16 == 0:
        col = x.reshape(n * h * w, c)
    elif c % 8 == 0:
        col = _Im2colFn.apply(x, k, k, s, p)
    else:   # stem (C = 3): no input gradient needed, plain im2col
        col = F.im2col(x, k, k, s, p)[0]
    y = matmul_fp8(col, conv, conv._w_bf16(), conv.k_true)
    return y.view(n, ho, wo, conv.out_channels)
