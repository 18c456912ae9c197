"""Functional wrappers over the sm_100a kernels (shape logic + dispatch).

Conventions
-----------
* activations are bf16, row-major; images are NHWC (``[N, H, W, C]``);
* a GEMM operand is a 2-D tensor with unit inner stride; ``*_mn=False`` means the
  tensor is ``[rows, K]`` (K-major), ``*_mn=True`` means it is ``[K, rows]``
  (MN-major) -- so forward / dgrad / wgrad need no transposes;
* kernels that produce parameter gradients *accumulate* into fp32 buffers (the
  flat gradient arena), the fused SGD step zeroes them again.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from ._ext import load

BF16 = torch.bfloat16
NUM_SMS = 148


def _pitch(t: torch.Tensor) -> int:
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), "GEMM operand must have unit inner stride"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _tma_ok(t: torch.Tensor) -> bool:
    return t.dtype == BF16 and _pitch(t) % 8 == 0 and t.data_ptr() % 16 == 0


def pick_bn(M: int, N: int) -> int:
    """Widest N tile that still yields >= one wave of CTAs; 64 for small problems."""
    m_tiles = (M + 127) // 128
    for bn in (256, 128):
        if N >= bn and m_tiles * ((N + bn - 1) // bn) >= NUM_SMS:
            return bn
    if N > 128 and m_tiles * ((N + 127) // 128) >= NUM_SMS // 2:
        return 128
    return 64


# measured on B200 (captured ResNet-18 step, 8 steps): 148 -> 4.538 ms, 96 -> 4.503, 64 -> 4.481, 32 -> 4.557
_WGRAD_MAX_CTAS = int(__import__("os").environ.get("BATON_WGRAD_MAX_CTAS", "64"))


def pick_split_k(M: int, N: int, K: int, bn: int) -> int:
    """Atomic split-K factor of a weight-gradient GEMM.  ``BATON_WGRAD_MAX_CTAS`` caps the CTAs of one launch: these
    GEMMs run as a parallel graph branch beside the dgrad / BatchNorm chain and should leave SMs to it."""
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    k_tiles = (K + 63) // 64
    if tiles >= NUM_SMS // 2 or k_tiles < 8:
        return 1
    return max(1, min(k_tiles // 4, max(1, _WGRAD_MAX_CTAS // tiles)))


def pick_cluster_k(M: int, N: int, K: int, bn: int) -> int:
    """Cluster split-K factor (1, 2, 4 or 8) for GEMMs with few output tiles and a long K."""
    import os
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    k_tiles = (K + 63) // 64
    min_kt = int(os.environ.get("BATON_GEMM_CLUSTER_MIN_KT", "4"))   # k tiles each CTA must keep
    if min_kt <= 0 or k_tiles < 16:     # short main loops gain nothing (8192x64x576: 5.5 us plain vs 6.7 us split)
        return 1
    best = 1
    for s in (2, 4, 8):
        # measured on B200: clusters of 8 only pay off while they cover at most ~half the SMs
        # (placement needs 8 free SMs inside one GPC); clusters of <= 4 are fine up to a full wave
        cap = NUM_SMS // 2 if s == 8 else 128
        if tiles * s <= cap and k_tiles >= min_kt * s:
            best = s
    return best


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = BF16, bias: Optional[torch.Tensor] = None,
         act: int = 0, accumulate: bool = False, alpha: float = 1.0, split_k: Optional[int] = None,
         n_valid: Optional[int] = None, flags: Optional[torch.Tensor] = None, flag_epoch: int = 0,
         flag_elem_off: int = 0, flag_tile_elems: int = 0, flag_bias_off: int = -1, force_bn: int = 0,
         force_simt: bool = False, col_stats: Optional[torch.Tensor] = None,
         flag_epoch_word: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[M,N] = act(alpha * A @ B^T + bias)`` on tcgen05 tensor cores.

    ``n_valid`` limits the written columns (used when B carries zero K-padding
    rows, e.g. the wgrad of a layer whose K was padded to a multiple of 8).

    ``col_stats`` (fp32 ``[2N]``): the epilogue also accumulates the per-column sum and sum of squares of the
    bf16 output into it -- the BatchNorm batch statistics of a convolution, without a second pass over the
    activation.  Only legal where :func:`gemm_stats_fusable` says so (single-pass tensor-core GEMM)."""
    C = load()
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, "GEMM reduction dims differ: {} vs {}".format(K, Kb)
    if n_valid is not None:
        N = min(N, n_valid)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
        if accumulate:
            out.zero_()
    ldd = out.stride(0) if out.dim() == 2 else N
    lda, ldb = _pitch(a), _pitch(b)
    use_simt = force_simt or not (_tma_ok(a) and _tma_ok(b))
    if use_simt:
        assert col_stats is None, "fused column statistics need the tensor-core path"
        C.gemm(a, b, out, bias, M, N, K, lda, ldb, ldd, a_mn, b_mn, act, 1, accumulate, alpha, None, 0, 0, 0, -1, 0,
               True, None, None)
        return out
    bn = force_bn or pick_bn(M, N)
    if col_stats is not None:
        assert not accumulate and bias is None and act == 0 and alpha == 1.0
    if split_k is None:
        if accumulate and out.dtype == torch.float32:
            split_k = pick_split_k(M, N, K, bn)          # atomic split-K straight into the gradient arena
        else:
            split_k = -pick_cluster_k(M, N, K, bn)       # cluster split-K, DSMEM reduce (negative = cluster)
            if split_k == -1:
                split_k = 1
    if split_k > 1:
        assert out.dtype == torch.float32 and bias is None and act == 0
    C.gemm(a, b, out, bias, M, N, K, lda, ldb, ldd, a_mn, b_mn, act, split_k, accumulate, alpha, flags, flag_epoch,
           flag_elem_off, flag_tile_elems, flag_bias_off, bn, False, col_stats, flag_epoch_word)
    return out


def gemm_stats_fusable(M: int, N: int, K: int) -> bool:
    """Can :func:`gemm` take column statistics of this problem in its epilogue?  Every bf16-output tensor-core
    path can: single-pass tiles do it from the TMEM rows (butterfly column sums), cluster split-K in the DSMEM
    reduction; only operands that fall back to the SIMT kernel (pitch not a multiple of 8) cannot."""
    return K % 8 == 0


# ---------------------------------------------------------------------------- elementwise / optimizer
def fused_sgd(w: torch.Tensor, g: torch.Tensor, hyper: torch.Tensor, momentum_buf: Optional[torch.Tensor] = None,
              w_bf16: Optional[torch.Tensor] = None, zero_grad: bool = True, nesterov: bool = False,
              max_ctas: int = 0, pack: Optional[dict] = None) -> None:
    """One kernel over the whole flat arena (reference: ``optimizer.step()``, demo.py:47).  ``max_ctas`` caps the
    grid for a slice that runs concurrently with other work.

    ``pack`` ("emits the upload copy"): ``{"wire_slot": int64[1] device word holding the wire address,
    "global_w": fp32 global copy or None, "scale": fp32[1] device scalar or None, "n_pack": elements to pack
    (parameters + float buffers), "wire_fp32": bool}`` -- the step also writes this client's wire copy for the
    round-end collective while the new weights are in registers."""
    pk = pack or {}
    load().fused_sgd(w, g, momentum_buf, w_bf16, hyper, zero_grad, nesterov, max_ctas, pk.get("wire_slot"),
                     pk.get("global_w"), pk.get("scale"), int(pk.get("n_pack", 0)), bool(pk.get("wire_fp32", False)))


def weighted_sum_(dst: torch.Tensor, srcs: Sequence[torch.Tensor], weights: Sequence[float]) -> torch.Tensor:
    """``dst = sum_k weights[k] * srcs[k]`` in one pass (manager-side FedAvg, manager.py:124-126)."""
    C = load()
    flat_dst = dst.view(-1) if dst.is_contiguous() else None
    if flat_dst is None or dst.dtype not in (torch.float32, BF16) or len(srcs) > C.MAX_RANKS:
        acc = torch.zeros_like(dst, dtype=torch.float32)
        for s, w in zip(srcs, weights):
            acc.add_(s.to(device=dst.device, dtype=torch.float32), alpha=float(w))
        dst.copy_(acc.to(dst.dtype))
        return dst
    ss = [s.to(device=dst.device, dtype=dst.dtype).contiguous().view(-1) for s in srcs]
    C.weighted_sum(flat_dst, ss, [float(w) for w in weights])
    return dst


def fold_client(acc: torch.Tensor, theta: torch.Tensor, global_w: torch.Tensor, nk: float, *, first: bool = False,
                reset: bool = False, w_bf16: Optional[torch.Tensor] = None, momentum: Optional[torch.Tensor] = None) -> None:
    """Time-sliced logical clients: ``acc (+)= nk * (theta - global)`` in one pass; ``reset`` also returns the replica
    to the global model (theta, bf16 shadow, momentum) for the next co-resident client."""
    load().fold_client(acc, theta, global_w, w_bf16, momentum, float(nk), 1 if first else 0, reset)


def fold_finish(acc: torch.Tensor, theta: torch.Tensor, global_w: torch.Tensor, total: float) -> None:
    """``theta = global + acc / total``: the sample-weighted mean replica this GPU uploads for its logical clients."""
    load().fold_client(acc, theta, global_w, None, None, 1.0 / float(total), 2, False)


def cast(src: torch.Tensor, dtype: torch.dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty(src.shape, dtype=dtype, device=src.device)
    load().cast(src.contiguous(), out)
    return out


def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``src[idx]`` for a resident on-GPU shard (reference: ``X[batch_idxs]``, demo.py:41-42)."""
    if out is None:
        out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    load().gather_rows(src, idx, out)
    return out


def colsum_(x2d: torch.Tensor, out: torch.Tensor, accumulate: bool = True) -> torch.Tensor:
    load().colsum(x2d, out, x2d.shape[0], x2d.shape[1], accumulate)
    return out


def add(a: torch.Tensor, b: torch.Tensor, relu: bool = False) -> torch.Tensor:
    out = torch.empty_like(a)
    load().add_bf16(a, b, out, relu)
    return out


def relu_bwd(y: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    dx = torch.empty_like(dy)
    load().relu_bwd(y, dy, dx)
    return dx


def gelu(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    load().gelu(x, y)
    return y


def gelu_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    dx = torch.empty_like(x)
    load().gelu_bwd(x, dy, dx)
    return dx


def pad_rows(src2d: torch.Tensor, kp: int, out: Optional[torch.Tensor] = None, gate: Optional[dict] = None) -> torch.Tensor:
    """``gate`` (bcast_gemm): ``{"flags", "epoch_word", "elem_off", "tile_elems"}`` -- wait for the FedAvg collective's
    arrival flags over the source slice of the bf16 arena before reading it."""
    rows, k = src2d.shape
    if out is None:
        out = torch.empty((rows, kp), dtype=BF16, device=src2d.device)
    g = gate or {}
    load().pad_rows(src2d, out, rows, k, kp, g.get("flags"), g.get("epoch_word"), int(g.get("elem_off", 0)),
                    int(g.get("tile_elems", 0)))
    return out


# ---------------------------------------------------------------------------- conv plumbing
def conv_out_size(h: int, k: int, stride: int, pad: int) -> int:
    return (h + 2 * pad - k) // stride + 1


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def im2col(x: torch.Tensor, kh: int, kw: int, stride: int, pad: int) -> Tuple[torch.Tensor, int, int, int]:
    """NHWC ``x`` -> ``col[N*Ho*Wo, Kp]`` with ``Kp = round_up(kh*kw*C, 8)``."""
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, kh, stride, pad), conv_out_size(w, kw, stride, pad)
    kp = round_up(kh * kw * c, 8)
    col = torch.empty((n * ho * wo, kp), dtype=BF16, device=x.device)
    load().im2col(x, col, n, h, w, c, kh, kw, stride, pad, ho, wo, kp)
    return col, ho, wo, kp


def conv_igemm_fwd(x: torch.Tensor, w2d: torch.Tensor, kh: int, kw: int, stride: int, pad: int,
                   col_stats: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """EXPERIMENTAL implicit-GEMM convolution forward: ``y[N*Ho*Wo, Cout]`` straight from NHWC ``x`` through TMA
    im2col loads (no ``col`` buffer).  ``w2d``: ``[Cout, kh*kw*Cin]`` channels_last weights.  Returns ``None``
    when the shape is not supported (Cin % 64 != 0)."""
    n, h, w, c = x.shape
    cout = w2d.shape[0]
    if c % 64 or w2d.shape[1] != kh * kw * c or not x.is_contiguous() or not w2d.is_contiguous():
        return None
    ho, wo = conv_out_size(h, kh, stride, pad), conv_out_size(w, kw, stride, pad)
    M, K = n * ho * wo, kh * kw * c
    bn = pick_bn(M, cout)
    y = torch.empty((M, cout), dtype=BF16, device=x.device)
    ok = load().conv_igemm_fwd(x, w2d, y, kh, kw, stride, pad, ho, wo, pick_cluster_k(M, cout, K, bn), bn, col_stats)
    return y if ok else None


def conv_igemm_dgrad(dy: torch.Tensor, w2d: torch.Tensor, in_shape, kh: int, kw: int, pad: int) -> Optional[torch.Tensor]:
    """Implicit-GEMM input gradient of a STRIDE-1 convolution: ``dx[N, H, W, Cin]`` from NHWC ``dy`` and the
    channels_last weights ``w2d [Cout, kh*kw*Cin]`` -- the flipped-filter convolution of ``dy``, gathered by TMA
    im2col, with the weight slab of each tap loaded MN-major in place (no ``dcol`` buffer, no col2im, no weight
    transpose).  Returns ``None`` when the shape is not supported (channels not multiples of 64)."""
    n, h, w, c = in_shape
    cout = dy.shape[-1]
    if c % 64 or cout % 64 or w2d.shape[1] != kh * kw * c or not dy.is_contiguous() or not w2d.is_contiguous():
        return None
    M, K = n * h * w, kh * kw * cout
    bn = pick_bn(M, c)
    dx = torch.empty((n, h, w, c), dtype=BF16, device=dy.device)
    ok = load().conv_igemm_dgrad(dy, w2d, dx, kh, kw, pad, pick_cluster_k(M, c, K, bn), bn)
    return dx if ok else None


def conv_igemm_wgrad_(dy2d: torch.Tensor, x: torch.Tensor, dw2d: torch.Tensor, kh: int, kw: int, stride: int,
                      pad: int) -> bool:
    """EXPERIMENTAL implicit wgrad: ``dw2d[Cout, kh*kw*Cin] += dy2d^T im2col(x)`` (fp32 atomics, split over
    pixels) without materialising ``im2col(x)``."""
    n, h, w, c = x.shape
    cout = dy2d.shape[1]
    if c % 64 or not x.is_contiguous() or not dy2d.is_contiguous() or not dw2d.is_contiguous():
        return False
    ho, wo = conv_out_size(h, kh, stride, pad), conv_out_size(w, kw, stride, pad)
    M, K = n * ho * wo, kh * kw * c
    bn = pick_bn(cout, K)
    return bool(load().conv_igemm_wgrad(dy2d, x, dw2d, cout, kh, kw, stride, pad, ho, wo, pick_split_k(cout, K, M, bn), bn))


def col2im(col: torch.Tensor, shape: Tuple[int, int, int, int], kh: int, kw: int, stride: int, pad: int,
           ho: int, wo: int) -> torch.Tensor:
    n, h, w, c = shape
    dx = torch.empty(shape, dtype=BF16, device=col.device)
    load().col2im(col, dx, n, h, w, c, kh, kw, stride, pad, ho, wo, col.shape[1])
    return dx


def maxpool(x: torch.Tensor, k: int, stride: int, pad: int) -> Tuple[torch.Tensor, torch.Tensor]:
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, k, stride, pad), conv_out_size(w, k, stride, pad)
    y = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    # winners: one byte (tap index inside the window) on the 16-byte path, a flat int32 position otherwise
    arg = torch.empty((n, ho, wo, c), dtype=torch.uint8 if (c % 8 == 0 and k * k <= 255) else torch.int32, device=x.device)
    load().maxpool(x, y, arg, n, h, w, c, k, stride, pad, ho, wo)
    return y, arg


def maxpool_bwd(dy: torch.Tensor, arg: torch.Tensor, in_shape, k: int, stride: int, pad: int,
                dy_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``dy_b``: optional second piece of the gradient (summed while loading; byte-argmax path only)."""
    n, h, w, c = in_shape
    if dy_b is not None and arg.dtype != torch.uint8:
        dy, dy_b = add(dy, dy_b), None
    dx = torch.empty(in_shape, dtype=BF16, device=dy.device)
    load().maxpool_bwd(dy, dy_b, arg, dx, n, h, w, c, dy.shape[1], dy.shape[2], k, stride, pad)
    return dx


def bn_relu_maxpool(z: torch.Tensor, sums: torch.Tensor, gamma, beta, rmean, rvar, nbt, eps: float, momentum: float,
                    k: int, stride: int, pad: int):
    """ResNet stem in one pass: ``maxpool(relu(batchnorm(z)))`` from the conv output ``z`` (NHWC bf16) and the batch
    statistic sums the conv GEMM's epilogue accumulated, without materialising the normalised activation.
    -> ``(p, argmax_u8, save_mean, save_rstd)`` or ``None`` when the shape is not supported (callers fall back to
    ``bn_apply`` + ``maxpool``).  Training mode only."""
    n, h, w, c = z.shape
    if c % 8 or k * k > 255:
        return None
    ho, wo = conv_out_size(h, k, stride, pad), conv_out_size(w, k, stride, pad)
    p = torch.empty((n, ho, wo, c), dtype=BF16, device=z.device)
    arg = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=z.device)
    mean = torch.empty(c, dtype=torch.float32, device=z.device)
    rstd = torch.empty(c, dtype=torch.float32, device=z.device)
    if not load().bn_relu_maxpool(z, p, arg, sums, gamma, beta, rmean, rvar, mean, rstd, nbt, n, h, w, c, k, stride, pad,
                                  ho, wo, eps, momentum):
        return None
    return p, arg, mean, rstd


def bn_maxpool_bwd(z: torch.Tensor, p: torch.Tensor, arg: torch.Tensor, dy_a: torch.Tensor, dy_b: Optional[torch.Tensor],
                   gamma, mean: torch.Tensor, rstd: torch.Tensor, sums_b: torch.Tensor, dgamma, dbeta, k: int, stride: int,
                   pad: int) -> Optional[torch.Tensor]:
    """Backward of :func:`bn_relu_maxpool`: ``dz`` from the pooled gradient ``dy_a (+ dy_b)``; ``dgamma`` / ``dbeta``
    are accumulated in place.  ``sums_b``: zeroed ``[2 * C]`` fp32 scratch.  ``None`` = shape not supported."""
    n, h, w, c = z.shape
    dz = torch.empty_like(z)
    if not load().bn_maxpool_bwd(z, p, arg, dy_a, dy_b, dz, gamma, mean, rstd, sums_b, dgamma, dbeta, n, h, w, c, k, stride,
                                 pad, p.shape[1], p.shape[2]):
        return None
    return dz


def avgpool(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    y = torch.empty((n, c), dtype=BF16, device=x.device)
    load().avgpool(x, y, n, h * w, c)
    return y


def avgpool_bwd(dy: torch.Tensor, in_shape) -> torch.Tensor:
    n, h, w, c = in_shape
    dx = torch.empty(in_shape, dtype=BF16, device=dy.device)
    load().avgpool_bwd(dy, dx, n, h * w, c)
    return dx


# ---------------------------------------------------------------------------- losses
def softmax_xent(logits: torch.Tensor, target: torch.Tensor, want_grad: bool = True,
                 grad_dtype: Optional[torch.dtype] = None, acc: Optional[torch.Tensor] = None):
    """Fused softmax cross-entropy: returns ``(acc, dlogits)`` where ``acc[0]`` is the
    batch-mean loss and ``acc[1]`` the number of correct predictions.  ``acc`` (fp32 ``[2]``) may be supplied: the
    kernel ADDS into it (a device-side running sum over the steps of an epoch, no extra kernels)."""
    rows, c = logits.shape
    if acc is None:
        acc = torch.zeros(2, dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits, dtype=grad_dtype or logits.dtype) if want_grad else None
    load().softmax_xent(logits, target, dl, acc, rows, c, logits.stride(0), 1.0 / rows)
    return acc, dl


def linear_xent_head(x: torch.Tensor, w_bf16: torch.Tensor, bias: Optional[torch.Tensor], target: torch.Tensor,
                     dw: torch.Tensor, db: Optional[torch.Tensor], acc: Optional[torch.Tensor] = None,
                     want_dx: bool = True, want_logits: bool = False):
    """Classifier head in one launch: ``logits = x w^T + b`` (<= 32 classes), softmax cross-entropy, and the head's whole
    backward -- ``dx`` (bf16), ``dw += dlogits^T x``, ``db += colsum(dlogits)`` (fp32, accumulated in place), the batch-mean
    loss / #correct added into ``acc``.  Returns ``(acc, dx, logits)`` or ``None`` when the shape is not supported."""
    rows, K = x.shape
    if acc is None:
        acc = torch.zeros(2, dtype=torch.float32, device=x.device)
    dx = torch.empty((rows, K), dtype=BF16, device=x.device) if want_dx else None
    logits = torch.empty((rows, w_bf16.shape[0]), dtype=torch.float32, device=x.device) if want_logits else None
    ok = load().linear_xent_head(x, w_bf16, bias, target, dx, dw, db, acc, logits, 1.0 / rows)
    return (acc, dx, logits) if ok else None


def mse(pred: torch.Tensor, target: torch.Tensor, want_grad: bool = True):
    acc = torch.zeros(2, dtype=torch.float32, device=pred.device)
    dp = torch.empty_like(pred) if want_grad else None
    load().mse(pred.contiguous(), target.contiguous().float(), dp, acc, 1.0 / pred.numel())
    return acc, dp


# ---------------------------------------------------------------------------- batched GEMM / attention / embedding
def gemm_batched(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldb: int,
                 ldd: int, a_mn: bool, b_mn: bool, n_outer: int, n_inner: int, a_strides, b_strides, d_strides,
                 alpha: float = 1.0, act: int = 0, accumulate: bool = False) -> torch.Tensor:
    """Strided-batched tcgen05 GEMM over a two-level batch (outer, inner) = (batch, head): operands are
    addressed through 4-D TMA tensor maps, so Q/K/V slices of a packed QKV buffer need no copies.
    ``*_strides`` = (outer, inner) element strides; ``a``/``b``/``out`` give the base pointers."""
    load().gemm_batched(a, b, out, M, N, K, lda, ldb, ldd, a_mn, b_mn, act, alpha, n_outer, n_inner,
                        a_strides[0], a_strides[1], b_strides[0], b_strides[1], d_strides[0], d_strides[1], accumulate)
    return out


def embedding_bwd_(dy2d: torch.Tensor, idx: torch.Tensor, grad_table: torch.Tensor) -> None:
    load().embedding_bwd(dy2d, idx, grad_table)


# ---------------------------------------------------------------------------- MXFP8 (block-scaled fp8)
def _sf_bytes(rows: int, k: int) -> int:
    return ((rows + 127) // 128) * ((k + 127) // 128) * 512


def quant_mx_rows(x2d: torch.Tensor):
    """bf16 ``[R, C]`` -> (e4m3 ``[R, Cp]`` as uint8, UE8M0 scale atoms); scales along C."""
    R, C = x2d.shape
    Cp = round_up(C, 16)
    q = torch.empty((R, Cp), dtype=torch.uint8, device=x2d.device)
    sf = torch.empty(_sf_bytes(R, C), dtype=torch.uint8, device=x2d.device)
    load().quant_mx_rows(x2d, q, sf, R, C, x2d.stride(0), Cp)
    return q, sf


def quant_mx_cols(x2d: torch.Tensor):
    """bf16 ``[R, C]`` -> (e4m3 ``[C, Rp]`` = quantised TRANSPOSE, scale atoms); scales along R."""
    R, C = x2d.shape
    Rp = round_up(R, 16)
    q = torch.zeros((C, Rp), dtype=torch.uint8, device=x2d.device) if Rp != R else \
        torch.empty((C, Rp), dtype=torch.uint8, device=x2d.device)
    sf = torch.empty(_sf_bytes(C, R), dtype=torch.uint8, device=x2d.device)
    load().quant_mx_cols(x2d, q, sf, R, C, x2d.stride(0), Rp)
    return q, sf


def dequant_mx(q: torch.Tensor, sf: torch.Tensor, cols: int) -> torch.Tensor:
    R, Cp = q.shape
    out = torch.empty((R, cols), dtype=torch.float32, device=q.device)
    load().dequant_mx(q, sf, out, R, cols, Cp)
    return out


def gemm_fp8(qa: torch.Tensor, sfa, qb: torch.Tensor, sfb, K: int, *, out: Optional[torch.Tensor] = None,
             out_dtype: torch.dtype = BF16, bias: Optional[torch.Tensor] = None, act: int = 0,
             accumulate: bool = False, alpha: float = 1.0, split_k: int = 1, n_valid: Optional[int] = None):
    """``out[M,N] = act(alpha * (A*SFA) @ (B*SFB)^T + bias)`` with e4m3 operands ``qa [M,Kp]``, ``qb [N,Kp]``."""
    M, N = qa.shape[0], qb.shape[0]
    if n_valid is not None:
        N = min(N, n_valid)
    if out is None:
        out = torch.zeros((M, N), dtype=out_dtype, device=qa.device) if accumulate else \
            torch.empty((M, N), dtype=out_dtype, device=qa.device)
    ldd = out.stride(0) if out.dim() == 2 else N
    load().gemm_fp8(qa, qb, out, bias, sfa, sfb, M, N, K, qa.stride(0), qb.stride(0), ldd, act, split_k, accumulate, alpha)
    return out
