"""torch.autograd Functions over the C ABI (include/taxo_b200.h): the only way compute reaches the GPU here.

Every Function below allocates its outputs/workspace with torch (device memory + stream plumbing) and calls the
hand-written CUDA kernels through ctypes; dense projections run on the tcgen05 3xTF32 kernels of tx_gemm.cu
(TAXO_GEMM=cublas switches them to torch.mm as a cross-check).  There is no eager/CPU fallback: a CPU tensor raises.

Feature matrices travel between layers as PADDED row-major buffers [N, ld] with ld % 4 == 0 (16-byte aligned rows
for 128-bit accesses / TMA); the logical width K <= ld is tracked by the caller, padding columns are zero.
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass
from typing import Optional

import torch
from torch.autograd import Function

from . import _lib
import ctypes

from ._lib import GatEpilogue, GatLayerDesc, GatLayerState, GcnLayerDesc, HeadDesc, HeadState, Stats, check, current_stream, device_guard, ptr, timed_region
from .graph import GraphStructure

_NEG_SLOPE_NONE = 1.0


def round4(k: int) -> int:
    return (int(k) + 3) // 4 * 4


def _check_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.TaxoLibraryError(f"{name} must be a CUDA tensor: taxoexpan_b200 has no CPU path (got device {t.device})")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (the reference path is fp32 end to end), got {t.dtype}")


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D view with unit column stride (copy only if needed)."""
    if t.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


FUSED_ENABLED = os.environ.get("TAXO_DISABLE_FUSED", "") == ""


FUSED_MAX_GRAPH_NODES = 2048
# apply the leaky-relu/dropout derivative of the previous layer's epilogue inside the d(z) GEMM epilogue, so the fused backward
# kernel reads d(z) as is (decoding the mask on each of its ~2.8 row loads cost 0.3 ms on L0; r11 / r13)
FUSE_DZ_EPILOGUE = os.environ.get("TAXO_FUSE_DZ_EPILOGUE", "1") not in ("", "0")
# the fused GAT kernels write their outputs already TF32-split (hi/lo) for the 3xTF32 GEMMs instead of a separate split pass
FUSE_SPLIT = os.environ.get("TAXO_FUSE_SPLIT", "1") not in ("", "0")
# TMA-staged fused GAT backward (tx_fused_bwd.cu) whenever d(z_next) needs no per-load mask decode; TAXO_STAGED_BWD=0 -> first kernel
STAGED_BWD = os.environ.get("TAXO_STAGED_BWD", "1") not in ("", "0")
# TMA-staged fused GAT forward (tx_fused_fwd.cu): parity-green but measured SLOWER than the warp-per-(row, head) kernel on the
# MAG-CS shapes (L0: 0.35 ms vs 0.23 ms - its one-thread-per-row softmax step and the two extra CTA barriers per tile cost more than
# the gather they save), so it is opt-in: TAXO_STAGED_FWD=1
STAGED_FWD = os.environ.get("TAXO_STAGED_FWD", "0") not in ("", "0")
# star-specialised fused GAT forward (tx_star_fwd.cu) for EgonetBatch structures: closed-form in-edges, every ft row loaded once per
# work item, anchor row in registers; TAXO_STAR_FWD=0 -> the general warp-per-(row, head) kernel
STAR_FWD = os.environ.get("TAXO_STAR_FWD", "1") not in ("", "0")
STAR_FWD_OUTPUT_LAYER = os.environ.get("TAXO_STAR_FWD", "1") != "hidden"     # TAXO_STAR_FWD=hidden: hidden layers only
# star-specialised fused GAT backward (tx_star_bwd.cu, second generation: warp-private TMA rings, (egonet, sibling chunk) work items,
# d(attn) through 2 H extra rows of the weight-gradient GEMM); TAXO_STAR_BWD=0 -> the tile-staged general backward
STAR_BWD = os.environ.get("TAXO_STAR_BWD", "1") not in ("", "0")
# the star backward first writes d(ft) with the fp16-pair scale of  TAXO_DFT_OPTIMISM x max|g| / (1 - p_attn)  and redoes the pass with
# the rigorous bound of tx_bound_dft only if a value left the fp16 range (device flag, no host sync); 0 = rigorous bound only
DFT_OPTIMISM = float(os.environ.get("TAXO_DFT_OPTIMISM", "4"))

_star_queues = {}


def _star_queue(device) -> torch.Tensor:
    """Work-queue counters of tx_gat_star_fwd (zero before the first launch, left zero by every launch): one buffer per
    (device, stream), because launches on different streams may overlap."""
    key = (device.index, current_stream().value)
    q = _star_queues.get(key)
    if q is None:
        lib = _lib.load()
        from .graph import STAR_MAX_CHUNKS
        if int(lib.tx_gat_star_max_chunks()) != STAR_MAX_CHUNKS:
            raise _lib.TaxoLibraryError("tx_gat_star_fwd: task encoding of the library differs from taxoexpan_b200.graph; rebuild")
        q = _star_queues[key] = torch.zeros(32 * 64, dtype=torch.int32, device=device)
    return q


# one C-ABI call per GAT layer and direction (tx_layer.cu: same kernels, arguments and order, one workspace) whenever the layer runs the
# default hot path (fp16-pair GEMMs, star forward / backward on an EgonetBatch); TAXO_LAYER_CALL=0 -> one ctypes call per kernel
LAYER_CALL = os.environ.get("TAXO_LAYER_CALL", "1") not in ("", "0")
# layer-0 input drop([x || P[pos]]) written directly as the fp16 operand pair of the first projection GEMM (tx_concat_pos_dropout_f16:
# one pass over x for max|x|, one pass that reads x and writes the pair) instead of fp32 z -> max|z| -> split (three passes over z);
# TAXO_CONCAT_F16=0 -> the fp32 z
CONCAT_F16 = os.environ.get("TAXO_CONCAT_F16", "1") not in ("", "0")
# parameter gradients written by the native backward calls directly into a registered flat gradient bucket (see grad_sink below);
# TAXO_GRAD_WRITE_THROUGH=0 -> freshly allocated gradients that autograd adds / copies into the bucket
GRAD_WRITE_THROUGH = os.environ.get("TAXO_GRAD_WRITE_THROUGH", "1") not in ("", "0")

_star_counter_bufs = {}
_star_rerun_bufs = {}


def _star_counters(device, n: int) -> torch.Tensor:
    """Per-egonet arrival counters of tx_gat_star_bwd (zero before the first launch, left zero by every launch), one growing buffer
    per (device, stream)."""
    key = (device.index, current_stream().value)
    c = _star_counter_bufs.get(key)
    if c is None or c.numel() < n:
        c = _star_counter_bufs[key] = torch.zeros(max(n, 1 << 16), dtype=torch.int32, device=device)
    return c


def star_bwd_reruns(device) -> torch.Tensor:
    """Device counter: how many tx_gat_star_bwd calls had to redo their pass with the rigorous fp16 scale (statistics)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    c = _star_rerun_bufs.get(idx)
    if c is None:
        c = _star_rerun_bufs[idx] = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", idx))
    return c


def use_fused(lib, heads: int, dim: int, mean_heads: int, st=None) -> bool:
    """Fused single-kernel GAT forward/backward when the shape qualifies (dim % 4 == 0, dim <= 512, heads == 1 for the
    head-mean output layer) and no graph of the batch is huge (the backward kernel gives one CTA per tile of whole
    graphs); otherwise the general-CSR kernels. TAXO_DISABLE_FUSED=1 forces the general path (tests)."""
    if st is not None and getattr(st, "max_nodes", 0) > FUSED_MAX_GRAPH_NODES:
        return False
    return FUSED_ENABLED and bool(lib.tx_gat_fused_supported(heads, dim, mean_heads))


# dense projections: "f16x3" = hand-written tcgen05 kernels of tx_gemm.cu on fp16 hi/lo operand pairs with per-tensor power-of-two
# scales (default: 22 significant bits per operand like 3xTF32, twice the tensor rate, half the operand bytes),
# "tf32x3" = the same kernels on TF32 hi/lo pairs, "cublas" = torch.mm (cuBLAS fp32 SIMT, TF32 off) kept as the yardstick
GEMM_BACKEND = os.environ.get("TAXO_GEMM", "f16x3")



def split_tf32(x: torch.Tensor, cols: int = None):
    """x [rows, >=cols] -> (hi, lo) padded [rows, round4(cols)] TF32-representable parts with x = hi + lo (+ 2^-22 rel.)."""
    lib = _lib.load()
    x = _rowmajor(x)
    rows = x.shape[0]
    cols = x.shape[1] if cols is None else cols
    ldo = round4(cols)
    hi = torch.empty((rows, ldo), dtype=torch.float32, device=x.device)
    lo = torch.empty((rows, ldo), dtype=torch.float32, device=x.device)
    check(lib.tx_split_tf32(ptr(x), x.stride(0) if rows > 1 else x.shape[1], rows, cols, ptr(hi), ptr(lo), ldo, current_stream()),
          "tx_split_tf32")
    return hi, lo


def gemm_nt_ps(a_hi, a_lo, k: int, b_hi, b_lo, n: int, out: torch.Tensor = None, epi=None) -> torch.Tensor:
    """C[:, :n] = A[:, :k] @ B[:n, :k]^T from pre-split operands on the tcgen05 3xTF32 kernel (optional fused output
    transform `epi`, a _lib.GemmEpilogue)."""
    lib = _lib.load()
    m = a_hi.shape[0]
    if out is None:
        out = torch.empty((m, round4(n)), dtype=torch.float32, device=a_hi.device)
    ldc = out.stride(0) if m > 1 else out.shape[1]
    if m > 0:
        check(lib.tx_gemm_nt_tf32x3_ex(ptr(a_hi), ptr(a_lo), a_hi.stride(0), ptr(b_hi), ptr(b_lo), b_hi.stride(0), ptr(out), ldc,
                                       m, n, k, epi, current_stream()), "tx_gemm_nt_tf32x3")
    return out if out.shape[1] == n else out[:, :n]


def gemm_tn_ps(a_hi, a_lo, m: int, b_hi, b_lo, n: int) -> torch.Tensor:
    """C[m, n] = sum_r A[r, :m]^T B[r, :n] (weight-gradient form) from pre-split operands; split-K + fixed-order reduce."""
    lib = _lib.load()
    r = a_hi.shape[0]
    ldc = round4(n)
    if r == 0:
        return torch.zeros((m, n), dtype=torch.float32, device=a_hi.device)
    splits = int(lib.tx_gemm_tn_splits(m, n, r))
    partial = torch.empty((splits, m, ldc), dtype=torch.float32, device=a_hi.device)
    check(lib.tx_gemm_tn_tf32x3(ptr(a_hi), ptr(a_lo), a_hi.stride(0), ptr(b_hi), ptr(b_lo), b_hi.stride(0), ptr(partial), ldc, m * ldc,
                                m, n, r, splits, current_stream()), "tx_gemm_tn_tf32x3")
    out = partial[0] if splits == 1 else _reduce_partials(lib, partial, splits, m * ldc).view(m, ldc)
    return out[:, :n]


def round8(n: int) -> int:
    return (n + 7) // 8 * 8


class F16Pair:
    """x * scale = hi + lo: the fp16 hi / lo parts [rows, round8(cols)] of an fp32 operand and its per-tensor power-of-two scale (a
    1-element DEVICE tensor - never read on the host).  `amax` (optional) is a device bound of max|x|."""
    __slots__ = ("hi", "lo", "scale", "cols")

    def __init__(self, hi, lo, scale, cols):
        self.hi, self.lo, self.scale, self.cols = hi, lo, scale, cols


def absmax(x: torch.Tensor, cols: int = None) -> torch.Tensor:
    """max |x[:, :cols]| as a 1-element device tensor (tx_absmax)."""
    lib = _lib.load()
    x = _rowmajor(x)
    rows = x.shape[0]
    cols = x.shape[1] if cols is None else cols
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    check(lib.tx_absmax(ptr(x), x.stride(0) if rows > 1 else x.shape[1], rows, cols, ptr(out), current_stream()), "tx_absmax")
    return out


def split_f16(x: torch.Tensor, cols: int = None, bound: torch.Tensor = None) -> F16Pair:
    """fp16 hi/lo split of x[:, :cols] (tx_split_f16).  `bound`: device upper bound of max|x| (default: measured with tx_absmax)."""
    lib = _lib.load()
    x = _rowmajor(x)
    rows = x.shape[0]
    cols = x.shape[1] if cols is None else cols
    if bound is None:
        bound = absmax(x, cols)
    ldo = round8(cols)
    hi = torch.empty((rows, ldo), dtype=torch.float16, device=x.device)
    lo = torch.empty((rows, ldo), dtype=torch.float16, device=x.device)
    scale = torch.empty(1, dtype=torch.float32, device=x.device)
    check(lib.tx_split_f16(ptr(x), x.stride(0) if rows > 1 else x.shape[1], rows, cols, ptr(bound), ptr(hi), ptr(lo), ldo, ptr(scale),
                           current_stream()), "tx_split_f16")
    return F16Pair(hi, lo, scale, cols)


def split_f16_weight(w: torch.Tensor):
    """One launch: max|w|, then the fp16 split of the weight matrix w [rows, cols] (row-major) in both orientations with the same
    scale.  Returns (F16Pair [rows, cols], F16Pair [cols, rows])."""
    lib = _lib.load()
    w = _rowmajor(w)
    rows, cols = w.shape
    ld, ldt = round8(cols), round8(rows)
    dev = w.device
    buf = torch.empty((rows * ld + cols * ldt) * 2, dtype=torch.float16, device=dev)      # hi | lo | hi_t | lo_t in one allocation
    hi, lo = buf[:rows * ld].view(rows, ld), buf[rows * ld:2 * rows * ld].view(rows, ld)
    o = 2 * rows * ld
    hit, lot = buf[o:o + cols * ldt].view(cols, ldt), buf[o + cols * ldt:].view(cols, ldt)
    scal = torch.empty(3, dtype=torch.float32, device=dev)                                # [amax, counter] workspace + scale
    scale = scal[2:3]
    check(lib.tx_split_f16_weight(ptr(w), w.stride(0) if rows > 1 else cols, rows, cols, ptr(hi), ptr(lo), ld, ptr(hit), ptr(lot), ldt,
                                  ptr(scal), ptr(scale), current_stream()), "tx_split_f16_weight")
    return F16Pair(hi, lo, scale, cols), F16Pair(hit, lot, scale, rows)


def gemm_nt_f16(a: F16Pair, k: int, b: F16Pair, n: int, out: torch.Tensor = None, epi=None, want_amax: bool = False):
    """C[:, :n] = A[:, :k] @ B[:n, :k]^T from fp16-split operands (tx_gemm_nt_f16x3).  Returns C, or (C, amax) with a device
    scalar max|C| when want_amax."""
    lib = _lib.load()
    m = a.hi.shape[0]
    if out is None:
        out = torch.empty((m, round4(n)), dtype=torch.float32, device=a.hi.device)
    ldc = out.stride(0) if m > 1 else out.shape[1]
    amax = torch.zeros(1, dtype=torch.float32, device=a.hi.device) if want_amax else None
    if m > 0:
        check(lib.tx_gemm_nt_f16x3(ptr(a.hi), ptr(a.lo), a.hi.stride(0), ptr(b.hi), ptr(b.lo), b.hi.stride(0), ptr(a.scale), ptr(b.scale),
                                   ptr(out), ldc, m, n, k, epi, ptr(amax), current_stream()), "tx_gemm_nt_f16x3")
    res = out if out.shape[1] == n else out[:, :n]
    return (res, amax) if want_amax else res


def gemm_tn_f16(a: F16Pair, m: int, b: F16Pair, n: int) -> torch.Tensor:
    """C[m, n] = sum_r A[r, :m]^T B[r, :n] from fp16-split operands; split-K + fixed-order reduce (tx_gemm_tn_f16x3)."""
    lib = _lib.load()
    r = a.hi.shape[0]
    ldc = round4(n)
    if r == 0:
        return torch.zeros((m, n), dtype=torch.float32, device=a.hi.device)
    splits = int(lib.tx_gemm_tn_f16_splits(m, n, r))
    partial = torch.empty((splits, m, ldc), dtype=torch.float32, device=a.hi.device)
    check(lib.tx_gemm_tn_f16x3(ptr(a.hi), ptr(a.lo), a.hi.stride(0), ptr(b.hi), ptr(b.lo), b.hi.stride(0), ptr(a.scale), ptr(b.scale),
                               ptr(partial), ldc, m * ldc, m, n, r, splits, current_stream()), "tx_gemm_tn_f16x3")
    out = partial[0] if splits == 1 else _reduce_partials(lib, partial, splits, m * ldc).view(m, ldc)
    return out[:, :n]


def gemm_nt(a: torch.Tensor, k: int, b: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """C = a[:, :k] @ b[:, :k]^T in fp32-faithful precision.  a: [M, >=k], b: [N, >=k] row-major.
    cublas backend: torch.mm (fp32 SIMT).  tf32x3 backend: split + tcgen05 3xTF32 kernel.
    `out` (optional) is a [M, ldc] buffer whose first N columns receive the result."""
    m, n = a.shape[0], b.shape[0]
    if GEMM_BACKEND == "f16x3" and m > 0:
        return gemm_nt_f16(split_f16(a, k), k, split_f16(b, k), n, out=out)
    if GEMM_BACKEND != "tf32x3" or m == 0:
        if out is None:
            return torch.mm(a[:, :k], b[:, :k].t())
        torch.mm(a[:, :k], b[:, :k].t(), out=out[:, :n])
        return out
    a_hi, a_lo = split_tf32(a, k)
    b_hi, b_lo = split_tf32(b, k)
    return gemm_nt_ps(a_hi, a_lo, k, b_hi, b_lo, n, out)


def _layer_gemms_fwd(z, k, w_nk, z_lo=None, z16=None, w_rowmajor=None, w_is_nk=True):
    """y = z[:, :k] @ w_nk[:, :k]^T.  Returns (y, saved, y_amax) where `saved` is what backward needs of z: z itself (cublas), its
    TF32 split (tf32x3) or its fp16 split (f16x3: (hi, lo, scale)) - the split is computed once and reused by the weight-gradient
    GEMM - and y_amax is a device scalar max|y| (f16x3 only: it bounds the next tensors' fp16 scales)."""
    if GEMM_BACKEND == "f16x3" and z.shape[0] > 0:
        zp = z16 if z16 is not None else split_f16(z, k)      # z16: produced pre-split by the previous layer's epilogue
        # the weights are split ONCE per step, in both orientations: [N, K] for this GEMM, [K, N] for the input-gradient GEMM
        if w_rowmajor is not None:
            p_a, p_b = split_f16_weight(w_rowmajor)
            wp, wt = (p_a, p_b) if w_is_nk else (p_b, p_a)
        else:
            wp, wt = split_f16(w_nk, k), None
        y, y_amax = gemm_nt_f16(zp, k, wp, w_nk.shape[0], want_amax=True)
        return y, (zp.hi, zp.lo, zp.scale, None if wt is None else wt.hi, None if wt is None else wt.lo, None if wt is None else wt.scale), y_amax
    if GEMM_BACKEND != "tf32x3" or z.shape[0] == 0:
        return torch.mm(z[:, :k], w_nk[:, :k].t()), (z, None, None, None, None, None), None
    if z_lo is None:
        z_hi, z_lo = split_tf32(z, k)
    else:
        z_hi = z                                      # produced pre-split by the previous layer's epilogue
    w_hi, w_lo = split_tf32(w_nk, k)
    return gemm_nt_ps(z_hi, z_lo, k, w_hi, w_lo, w_nk.shape[0]), (z_hi, z_lo, None, None, None, None), None


def _dz_epilogue(in_link, c0a):
    if FUSE_DZ_EPILOGUE and in_link is not None and in_link.mask is not None and c0a == 0:
        return _lib.GemmEpilogue(act_mask=ptr(in_link.mask), heads=in_link.heads, dim=in_link.dim,
                                 mask_stride=int(_lib.load().tx_gat_fused_mask_ld(in_link.heads, in_link.dim)),
                                 col0=0, act_slope=in_link.act_slope, p_drop=in_link.p_drop,
                                 has_keep_plane=1 if in_link.p_drop > 0.0 else 0)
    return None


def _layer_gemms_bwd_f16(dy, f, saved, w_kf, k, ldz, c0, need_w, need_z, in_link=None, dy16=None, extra_rows=0):
    """f16x3 form of _layer_gemms_bwd: dy comes fp16-split from the fused backward kernel (dy16) or is split here.  extra_rows: columns
    [f, f + extra_rows) of dy16 (the star backward's per-node attention coefficients) join the weight-gradient GEMM: dw has
    f + extra_rows rows."""
    n = saved[0].shape[0]
    dw = dz = None
    zp = F16Pair(saved[0], saved[1], saved[2], k)
    if dy16 is None:
        with timed_region("split_dy"):
            dy16 = split_f16(dy, f)
    if need_w:
        with timed_region("gemm_dw"):
            dw = gemm_tn_f16(dy16, f + extra_rows, zp, k)
    if need_z:
        with timed_region("gemm_dz"):
            c0a = (min(c0, k) // 8) * 8                   # fp16 rows: 16-byte aligned slices start at multiples of 8 columns
            dz = torch.empty((n, ldz), dtype=torch.float32, device=zp.hi.device)
            if k > c0a:
                if len(saved) > 3 and saved[3] is not None:       # [K, F] orientation split in the forward pass
                    wp = F16Pair(saved[3][c0a:k], saved[4][c0a:k], saved[5], f)
                else:
                    wp = split_f16(w_kf[c0a:k], f)
                epi = _dz_epilogue(in_link, c0a)
                _, dz_amax = gemm_nt_f16(dy16, f, wp, k - c0a, out=dz[:, c0a:], epi=epi, want_amax=True)
                if in_link is not None:
                    in_link.dz_amax = dz_amax             # bounds |g| of the layer below (its fused backward writes fp16-split dft)
                    if epi is not None:
                        in_link.applied = True
            if ldz > k and round4(k) < ldz:
                dz[:, round4(k):].zero_()
    return dw, dz


def _layer_gemms_bwd(dy, f, saved, w_kf, k, ldz, c0, need_w, need_z, in_link=None, dy_lo=None, dy16=None, extra_rows=0):
    """dW_fk = dy[:, :f]^T @ z[:, :k]  and  dz[:, c0a:k] = dy[:, :f] @ w_kf[c0a:k, :f]^T (see _gemm_dz)."""
    if GEMM_BACKEND == "f16x3" and saved[0].shape[0] > 0:
        return _layer_gemms_bwd_f16(dy, f, saved, w_kf, k, ldz, c0, need_w, need_z, in_link, dy16, extra_rows)
    n = dy.shape[0]
    dw = dz = None
    if GEMM_BACKEND != "tf32x3" or n == 0:
        z = saved[0]
        if need_w:
            with timed_region("gemm_dw"):
                dw = torch.mm(dy[:, :f].t(), z[:, :k])
        if need_z:
            with timed_region("gemm_dz"):
                c0a = (min(c0, k) // 4) * 4
                dz = torch.empty((n, ldz), dtype=torch.float32, device=dy.device)
                if k > c0a:
                    torch.mm(dy[:, :f], w_kf[c0a:k, :f].t(), out=dz[:, c0a:k])
                if ldz > k:
                    dz[:, k:].zero_()
        return dw, dz
    z_hi, z_lo = saved[0], saved[1]
    if dy_lo is not None:
        d_hi, d_lo = dy, dy_lo                        # written pre-split by the fused backward kernel
    else:
        with timed_region("split_dy"):
            d_hi, d_lo = split_tf32(dy, f)
    if need_w:
        with timed_region("gemm_dw"):
            dw = gemm_tn_ps(d_hi, d_lo, f, z_hi, z_lo, k)
    if need_z:
        with timed_region("gemm_dz"):
            c0a = (min(c0, k) // 4) * 4
            dz = torch.empty((n, ldz), dtype=torch.float32, device=dy.device)
            if k > c0a:
                w_hi, w_lo = split_tf32(w_kf[c0a:k], f)
                epi = _dz_epilogue(in_link, c0a)
                gemm_nt_ps(d_hi, d_lo, f, w_hi, w_lo, k - c0a, out=dz[:, c0a:], epi=epi)
                if epi is not None:
                    in_link.applied = True
    if os.environ.get("TAXO_DEBUG_GEMM") and in_link is None:
        torch.cuda.synchronize()
        z = (z_hi + z_lo)
        print(f"[debug gemm] f={f} k={k} n={n} dy finite {bool(torch.isfinite(dy).all())} split err {float((d_hi + d_lo - dy[:, :f]).abs().max()):.2e}")
        if dw is not None:
            ref = torch.mm(dy[:, :f].t(), z[:, :k])
            d = (dw - ref).abs()
            print(f"   dw diff {float(d.max()):.3e} at {divmod(int(d.argmax()), k)} ref {float(ref.abs().max()):.3e}")
        if dz is not None:
            ref = torch.mm(dy[:, :f], w_kf[:k, :f].t())
            d = (dz[:, :k] - ref).abs()
            print(f"   dz diff {float(d.max()):.3e} at {divmod(int(d.argmax()), k)} ref {float(ref.abs().max()):.3e}")
    return dw, dz


def new_seed() -> int:
    """A fresh 62-bit dropout seed from torch's CPU generator (so torch.manual_seed makes runs repeatable)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _reduce_partials(lib, partial: torch.Tensor, n_blocks: int, m_len: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty(m_len, dtype=torch.float32, device=partial.device)
    check(lib.tx_reduce_partials(ptr(partial), n_blocks, m_len, ptr(out), current_stream()), "tx_reduce_partials")
    return out


def gather_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """x[i] = table[ids[i]]: the feature rows of a batch from the node-embedding table kept RESIDENT on the device (tx_gather_rows), so
    a training step ships node ids instead of rows (the reference collates rows of g_full.ndata['x'] on the host per egonet,
    data_loader/dataset.py:157,429-431).  ids: int32 / int64 device tensor; the table carries no gradient (trainer.py:48 never asks
    for one)."""
    lib = _lib.load()
    _check_cuda(table, "feature table")
    if table.dim() != 2 or table.stride(1) != 1:
        raise ValueError("gather_rows: the table must be a row-major 2-D tensor")
    if ids.device != table.device:
        raise ValueError("gather_rows: ids must live on the table's device")
    ids32 = ids if ids.dtype == torch.int32 else ids.to(torch.int32)
    ids32 = ids32.contiguous()
    n, d = int(ids32.numel()), int(table.shape[1])
    out = torch.empty((n, d), dtype=torch.float32, device=table.device)
    if n == 0:
        return out
    with device_guard(table.device):
        check(lib.tx_gather_rows(ptr(table), table.stride(0), table.shape[0], ptr(ids32), n, d, ptr(out), d, current_stream()), "tx_gather_rows")
    return out


def dropout_keep_mask(seed: int, stream_id: int, first_index: int, n: int, p: float, device) -> torch.Tensor:
    """The exact keep-mask (uint8, 1 = keep) the kernels use; lets the CPU oracle replay a dropout run."""
    lib = _lib.load()
    keep = torch.empty(n, dtype=torch.uint8, device=device)
    with device_guard(keep.device):
        check(lib.tx_dropout_keep_mask(seed, stream_id, first_index, n, p, ptr(keep), current_stream()), "tx_dropout_keep_mask")
    return keep


# --------------------------------------------------------------------------------------------------
# z = drop([x || P[pos]])
# --------------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------------
# Gradient sinks: a flat gradient bucket (taxoexpan_b200/dist.py) registers, per parameter, the slice of its buffer that is the
# parameter's .grad.  The native backward calls then write a gradient straight into that slice and return an alias of it, so
# autograd's AccumulateGrad adopts the tensor as .grad without an add / copy kernel (it "steals" a contiguous gradient nobody else
# references when .grad is None).  Only when the parameter's .grad is None - with an existing .grad autograd must accumulate, and the
# usual freshly allocated gradient is returned.
# --------------------------------------------------------------------------------------------------
_grad_sinks = {}


def register_grad_sink(param: torch.Tensor, flat_view: torch.Tensor):
    _grad_sinks[param.data_ptr()] = (weakref.ref(param), flat_view)


def unregister_grad_sink(param: torch.Tensor):
    _grad_sinks.pop(param.data_ptr(), None)


def grad_sink(param_like: Optional[torch.Tensor], numel: int) -> Optional[torch.Tensor]:
    """A fresh 1-D alias of the registered home of this parameter's gradient (matched by data pointer, so views like W.weight.view(l, r)
    find their parameter), or None when there is none / the parameter already holds a gradient that autograd has to add to."""
    if param_like is None or not _grad_sinks or not GRAD_WRITE_THROUGH:
        return None
    ent = _grad_sinks.get(param_like.data_ptr())
    if ent is None:
        return None
    p = ent[0]()
    if p is None:
        _grad_sinks.pop(param_like.data_ptr(), None)
        return None
    v = ent[1]
    if p.grad is not None or v.numel() != numel or v.device != param_like.device or not v.is_contiguous():
        return None
    return v.view(-1)


def concat_publishes_f16() -> bool:
    """The layer-0 input is written straight as the fp16 operand pair of the first projection GEMM (tx_concat_pos_dropout_f16) when the
    default dense back-end consumes such pairs; any other configuration gets the fp32 z of tx_concat_pos_dropout_fwd."""
    return CONCAT_F16 and GEMM_BACKEND == "f16x3" and FUSE_SPLIT


class ConcatPosDropout(Function):
    @staticmethod
    def forward(ctx, x, pos_table, pos32, p, seed, stream_id, link=None):
        lib = _lib.load()
        _check_cuda(x, "features")
        x = _rowmajor(x)
        n, k_in = x.shape
        pd = 0 if pos_table is None else int(pos_table.shape[1])
        vocab = 0 if pos_table is None else int(pos_table.shape[0])
        ldz = round4(k_in + pd)
        z = torch.empty((n, ldz), dtype=torch.float32, device=x.device)
        tab = None if pos_table is None else pos_table.contiguous()
        if link is not None and n > 0 and concat_publishes_f16():
            # z is published to the first layer as an fp16 hi / lo pair in one workspace [hi | lo | scale]; `z` stays a placeholder
            # (never written), exactly like the output of a hidden layer that hands its successor a pair
            ld16 = round8(k_in + pd)
            half_bytes = n * ld16 * 2
            with device_guard(x.device):
                Stats.tag = "L0"
                ws = torch.empty(2 * half_bytes + 16, dtype=torch.uint8, device=x.device)
                x_amax = absmax(x, k_in)
                base = ws.data_ptr()
                check(lib.tx_concat_pos_dropout_f16(ptr(x), x.stride(0) if n > 1 else k_in, ptr(tab), ptr(pos32), n, k_in, pd, vocab, p, seed,
                                                    stream_id, ptr(x_amax), base, base + half_bytes, ld16, base + 2 * half_bytes,
                                                    current_stream()), "tx_concat_pos_dropout_f16")
            state = _lib.GatLayerState()
            state.out_hi, state.out_lo, state.out_scale, state.ld16_out = base, base + half_bytes, base + 2 * half_bytes, ld16
            link.c_state, link.c_ws, link.c_dims = state, ws, (n, k_in + pd, ld16, 0)
            link.applied, link.z16, link.z_lo, link.dz_amax, link.c_bwd, link.mask = False, None, None, None, None, None
            ctx.save_for_backward(pos32 if pos32 is not None else torch.empty(0))
            ctx.meta = (n, k_in, pd, vocab, ldz, float(p), int(seed), int(stream_id))
            ctx.tab_ref = pos_table
            return z
        with device_guard(x.device):
            check(lib.tx_concat_pos_dropout_fwd(ptr(x), x.stride(0) if n > 1 else k_in, ptr(tab), ptr(pos32), n, k_in, pd,
                                                ptr(z), ldz, p, seed, stream_id, current_stream()),
                  "tx_concat_pos_dropout_fwd")
        ctx.save_for_backward(pos32 if pos32 is not None else torch.empty(0))
        ctx.meta = (n, k_in, pd, vocab, ldz, float(p), int(seed), int(stream_id))
        ctx.tab_ref = pos_table
        return z

    @staticmethod
    def backward(ctx, dz):
        lib = _lib.load()
        (pos32,) = ctx.saved_tensors
        n, k_in, pd, vocab, ldz, p, seed, stream_id = ctx.meta
        need_x, need_tab = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and pd > 0
        dx = dtab = None
        if not (need_x or need_tab):
            return None, None, None, None, None, None, None
        dz = _rowmajor(dz)
        if n > 1 and dz.stride(0) != ldz:
            dz = dz.contiguous()              # the keep-mask counters are indexed with the forward's row pitch ldz
        if not need_x:
            # x carries no gradient (the usual case, trainer.py:48): only d(position table) = sum over rows of the kept,
            # rescaled position columns is needed - one read of [N, pos_dim] instead of a rescale of the whole d(z)
            with device_guard(dz.device):
                nb = int(lib.tx_row_blocks(n))
                partial = torch.empty(nb * vocab * pd, dtype=torch.float32, device=dz.device)
                check(lib.tx_pos_grad_partials(ptr(dz), dz.stride(0) if n > 1 else ldz, k_in, ptr(pos32), n, pd, vocab, p, seed, stream_id,
                                               ptr(partial), current_stream()), "tx_pos_grad_partials")
                dtab = _reduce_partials(lib, partial, nb, vocab * pd, out=grad_sink(ctx.tab_ref, vocab * pd)).view(vocab, pd)
            return None, dtab, None, None, None, None, None
        if p > 0.0:
            dz = dz.clone()      # the kernel rescales the kept entries in place; never touch the caller's grad
        with device_guard(dz.device):
            nb = int(lib.tx_row_blocks(n))
            partial = torch.empty(nb * vocab * pd, dtype=torch.float32, device=dz.device) if need_tab else None
            # with p == 0 and no activation the kernel's feature pass is a no-op (dz is only read)
            check(lib.tx_epilogue_bwd(ptr(dz), dz.stride(0) if n > 1 else ldz, None, ptr(pos32) if pd > 0 else None, n, k_in,
                                      pd if need_tab else 0, vocab, 1.0, p, seed, stream_id, ptr(partial), current_stream()),
                  "tx_epilogue_bwd")
            if need_tab:
                dtab = _reduce_partials(lib, partial, nb, vocab * pd).view(vocab, pd)
        if need_x:
            dx = dz[:, :k_in].contiguous()
        return dx, dtab, None, None, None, None, None


# --------------------------------------------------------------------------------------------------
# GAT layer: ft = z W^T ; fused attention/softmax/aggregate ; epilogue = next layer's input or head mean
# --------------------------------------------------------------------------------------------------
class MaskLink:
    """Hand-shake between consecutive fused GAT layers: layer l-1's forward publishes the sign/keep bytes of its epilogue,
    layer l's backward applies their derivative inside its d(z) GEMM epilogue and flags it, so layer l-1's backward kernel
    reads d(z) as is (no per-load decode)."""
    __slots__ = ("mask", "heads", "dim", "act_slope", "p_drop", "applied", "z_lo", "z16", "dz_amax", "c_state", "c_ws", "c_dims", "c_bwd")

    def __init__(self):
        self.mask = None
        self.applied = False
        self.z_lo = None      # TF32 "lo" part of the published z when the producer wrote z pre-split (z itself is then the "hi" part)
        self.z16 = None       # F16Pair of the published z when the producer wrote it fp16-split (z itself is then never written)
        self.dz_amax = None   # device scalar max|d(z)| published by the consumer layer's backward GEMM epilogue
        self.c_state = None   # tx_gat_layer_state of a producer that ran through tx_gat_layer_fwd (pointers into c_ws)
        self.c_ws = None
        self.c_dims = None    # (n, cols, ld16, mask words) of the published pair / mask
        self.c_bwd = None     # (device pointer of max|d(z)|, backward workspace keeping it alive) published by a native consumer

    def materialize(self):
        """Tensor views (z16, mask) of what a native producer published, for a consumer on the per-kernel path."""
        if self.c_state is None or self.z16 is not None:
            return
        n, cols, ld16, words = self.c_dims
        ws, s_ = self.c_ws, self.c_state
        base = ws.data_ptr()

        def view(p, nbytes, dtype):
            return ws[p - base:p - base + nbytes].view(dtype)
        self.z16 = F16Pair(view(s_.out_hi, n * ld16 * 2, torch.float16).view(n, ld16), view(s_.out_lo, n * ld16 * 2, torch.float16).view(n, ld16),
                           view(s_.out_scale, 4, torch.float32), cols)
        if s_.maskbits:
            self.mask = view(s_.maskbits, words * 4, torch.int32)


@dataclass
class GatLayerCfg:
    k: int                      # logical input width (z[:, :k])
    heads: int
    dim: int                    # per-head width D'
    neg_slope: float = 0.2      # leaky-relu slope inside attention (model_zoo.py:70)
    p_attn: float = 0.0
    attn_seed: int = 0
    attn_stream: int = 1
    hidden: bool = True         # True: emit next layer's input; False: output layer (mean over heads)
    act_slope: float = 1.0      # activation after a hidden layer (0.01 = F.leaky_relu default); 1 = none
    p_next: float = 0.0         # next layer's feat_drop, applied by this layer's epilogue
    next_seed: int = 0
    next_stream: int = 0
    dz_from: int = 0            # columns [0, dz_from) of d(z) are not needed by the caller (layer 0, x without grad)
    tag: str = ""               # label for bench.py's per-kernel CUDA-event timings
    in_link: Optional[MaskLink] = None    # published by the previous layer (its epilogue produced this layer's z)
    out_link: Optional[MaskLink] = None   # published to the next layer


def _native_layer_ok(lib, z, weight, attn_l, attn_r, next_pos_table, st, cfg, n) -> bool:
    """The layer runs the default hot path end to end: one native call per direction (tx_gat_layer_fwd / _bwd)."""
    if not (LAYER_CALL and GEMM_BACKEND == "f16x3" and FUSE_SPLIT and FUSE_DZ_EPILOGUE and STAR_FWD and STAR_BWD and STAGED_BWD
            and FUSED_ENABLED and not STAGED_FWD and n > 0):
        return False
    if st.star is None or st.star_bwd is None or cfg.heads > 64 or not use_fused(lib, cfg.heads, cfg.dim, 0 if cfg.hidden else 1, st):
        return False
    if cfg.hidden and cfg.out_link is None:
        return False
    if not cfg.hidden and not STAR_FWD_OUTPUT_LAYER:
        return False
    if weight.shape[1] != cfg.k or not weight.is_contiguous() or not attn_l.is_contiguous() or not attn_r.is_contiguous():
        return False
    if next_pos_table is not None and not next_pos_table.is_contiguous():
        return False
    if cfg.in_link is not None and cfg.in_link.c_state is None:
        return False                                    # the layer below published tensors, not a native state
    if cfg.in_link is None and (z.stride(1) != 1 or z.stride(0) % 4 != 0 or z.data_ptr() % 16 != 0):
        return False
    return True


def _native_desc(lib, st, cfg, n, weight, attn_l, attn_r, next_pos_table, pos32, dev):
    H, D = cfg.heads, cfg.dim
    d = GatLayerDesc()
    d.n, d.e, d.k, d.heads, d.dim = n, st.e, cfg.k, H, D
    pd = 0 if (next_pos_table is None or not cfg.hidden) else int(next_pos_table.shape[1])
    d.pos_dim, d.vocab = pd, (0 if pd == 0 else int(next_pos_table.shape[0]))
    d.dz_from, d.max_out_deg, d.hidden = cfg.dz_from, max(int(st.max_out_deg), 1), 1 if cfg.hidden else 0
    d.neg_slope, d.p_attn, d.act_slope = cfg.neg_slope, cfg.p_attn, cfg.act_slope
    d.p_next, d.dft_optimism = (cfg.p_next if cfg.hidden else 0.0), DFT_OPTIMISM
    d.attn_seed, d.next_seed, d.attn_stream, d.next_stream = cfg.attn_seed, cfg.next_seed, cfg.attn_stream, cfg.next_stream
    sf, sb = st.star, st.star_bwd
    d.tasks_fwd, d.n_tasks_fwd, d.chunk_fwd = sf[0].data_ptr(), sf[1], sf[2]
    d.tasks_bwd, d.n_tasks_bwd, d.chunk_bwd = sb[0].data_ptr(), sb[1], sb[2]
    d.pos = None if pd == 0 else pos32.data_ptr()
    d.queue = _star_queue(dev).data_ptr()
    d.counters = _star_counters(dev, sb[1] * H).data_ptr()
    d.reruns = star_bwd_reruns(dev).data_ptr()
    d.weight, d.ldw = weight.data_ptr(), weight.stride(0)
    d.attn_l, d.attn_r = attn_l.data_ptr(), attn_r.data_ptr()
    d.next_pos_table = None if pd == 0 else next_pos_table.data_ptr()
    d.tag = cfg.tag.encode()[:15]
    return d, pd


class GatLayer(Function):
    @staticmethod
    def forward(ctx, z, weight, attn_l, attn_r, next_pos_table, st: GraphStructure, pos32, cfg: GatLayerCfg):
        lib = _lib.load()
        _check_cuda(z, "z")
        n, ldz = z.shape
        H, D, K = cfg.heads, cfg.dim, cfg.k
        F_ = H * D
        dev = z.device
        f32 = dict(dtype=torch.float32, device=dev)
        ctx.native = None
        if _native_layer_ok(lib, z, weight, attn_l, attn_r, next_pos_table, st, cfg, n):
            # ---- one native call: split / GEMM / bound / star forward into one workspace (tx_layer.cu) ----
            with device_guard(dev):
                Stats.sync_native_profiling()
                desc, pd = _native_desc(lib, st, cfg, n, weight, attn_l, attn_r, next_pos_table, pos32, dev)
                prev = None if cfg.in_link is None else cfg.in_link.c_state
                ws = torch.empty(int(lib.tx_gat_layer_fwd_bytes(ctypes.byref(desc), 1 if prev is None else 0)), dtype=torch.uint8, device=dev)
                state = GatLayerState()
                out = torch.empty((n, round4(F_ + pd)) if cfg.hidden else (n, D), **f32)     # hidden: a placeholder, never written
                check(lib.tx_gat_layer_fwd(ctypes.byref(desc), ptr(z) if prev is None else None, ldz,
                                           None if prev is None else ctypes.byref(prev), ptr(ws), ctypes.byref(state),
                                           None if cfg.hidden else ptr(out), current_stream()), "tx_gat_layer_fwd")
            if cfg.out_link is not None:
                lk = cfg.out_link
                lk.c_state, lk.c_ws, lk.applied, lk.z16, lk.z_lo, lk.dz_amax, lk.c_bwd = state, ws, False, None, None, None, None
                lk.c_dims = (n, F_ + pd, round8(F_ + pd), int(lib.tx_gat_fused_mask_words(n, H, D)))
                lk.mask = None
                if state.maskbits:
                    lk.heads, lk.dim, lk.act_slope, lk.p_drop = H, D, cfg.act_slope, cfg.p_next
            ctx.native = (desc, state, ws, pd)
            ctx.fused, ctx.maskbits = True, None
            ctx.st, ctx.cfg, ctx.pd = st, cfg, pd
            ctx.vocab = 0 if next_pos_table is None else int(next_pos_table.shape[0])
            ctx.save_for_backward(weight, attn_l, attn_r, next_pos_table, pos32, z if prev is None else None)
            ctx.attn_shape = attn_l.shape
            ctx.zshape = tuple(z.shape)
            return out
        if cfg.in_link is not None:
            cfg.in_link.materialize()
        with device_guard(dev):
            stream = current_stream()
            Stats.tag = cfg.tag
            f16 = GEMM_BACKEND == "f16x3"
            with timed_region("gemm_fwd"):
                ft, zsaved, ft_amax = _layer_gemms_fwd(z, K, weight,      # ft = fc(h), model_zoo.py:83
                                                       cfg.in_link.z_lo if (cfg.in_link is not None and GEMM_BACKEND == "tf32x3") else None,
                                                       cfg.in_link.z16 if (cfg.in_link is not None and f16) else None,
                                                       w_rowmajor=weight if (f16 and weight.shape[1] == K) else None)
            al = attn_l.reshape(-1).contiguous()
            ar = attn_r.reshape(-1).contiguous()
            alpha = torch.empty(st.e * H, **f32)
            elog = torch.empty(st.e * H, **f32)
            alpha_d = torch.empty(st.e * H, **f32) if cfg.p_attn > 0.0 else alpha
            pd = 0 if next_pos_table is None else int(next_pos_table.shape[1])
            tab = None if next_pos_table is None else next_pos_table.contiguous()
            if cfg.hidden:
                ldo = round4(F_ + pd)
                out = torch.empty((n, ldo), **f32)
            else:
                ldo = D
                out = torch.empty((n, D), **f32)
            epi = GatEpilogue(mean_heads=0 if cfg.hidden else 1, act_slope=cfg.act_slope, next_pos_table=ptr(tab),
                              pos=ptr(pos32) if pd > 0 else None, pos_dim=pd, p_drop=cfg.p_next if cfg.hidden else 0.0,
                              seed=cfg.next_seed, stream_id=cfg.next_stream)
            fused = use_fused(lib, H, D, 0 if cfg.hidden else 1, st)
            star = fused and STAR_FWD and st.star is not None and H <= 64 and n > 0
            maskbits = None
            out_lo = None
            out16 = None
            if FUSE_SPLIT and fused and cfg.hidden and cfg.out_link is not None and GEMM_BACKEND == "tf32x3":
                out_lo = torch.empty_like(out)        # the epilogue writes the next layer's input already TF32-split
            if fused:
                # ONE kernel: logits from the gathered rows, edge softmax, dropout, aggregation, next-layer epilogue
                if cfg.hidden and (cfg.act_slope != 1.0 or cfg.p_next > 0.0):
                    maskbits = torch.empty(int(lib.tx_gat_fused_mask_words(n, H, D)), dtype=torch.int32, device=dev)
                if FUSE_SPLIT and f16 and cfg.hidden and cfg.out_link is not None and ft_amax is not None and n > 0:
                    # the epilogue writes the next layer's input fp16-split (x * scale = hi + lo); `out` itself is never written.
                    # |z_next| <= max|ft| / ((1 - p_attn)(1 - p_next)) (attention weights are convex), appended rows <= max|P| / (1 - p_next)
                    ld16 = round8(F_ + pd)
                    o_hi = torch.empty((n, ld16), dtype=torch.float16, device=dev)
                    o_lo = torch.empty((n, ld16), dtype=torch.float16, device=dev)
                    o_scale = torch.empty(1, **f32)
                    bound = torch.empty(1, **f32)
                    check(lib.tx_bound_max2(ptr(ft_amax), 1.0 / ((1.0 - cfg.p_attn) * (1.0 - cfg.p_next)), ptr(tab) if pd > 0 else None,
                                            tab.numel() if pd > 0 else 0, 1.0 / (1.0 - cfg.p_next), ptr(bound), stream), "tx_bound_max2")
                    if star:
                        sg = st.star
                        check(lib.tx_gat_star_fwd(ptr(ft), F_, ptr(al), ptr(ar), ptr(sg[0]), sg[1], sg[2], n, H, D, cfg.neg_slope, cfg.p_attn, cfg.attn_seed, cfg.attn_stream, ptr(alpha),
                                                  ptr(alpha_d), ptr(elog), None, ldo, epi, ptr(maskbits), ptr(o_hi), ptr(o_lo), ld16,
                                                  ptr(bound), ptr(o_scale), ptr(_star_queue(dev)), stream), "tx_gat_star_fwd")
                    elif STAGED_FWD:
                        check(lib.tx_gat_fused_fwd_staged(ptr(ft), F_, ptr(al), ptr(ar), ptr(st.in_ptr), ptr(st.in_src), ptr(st.in_eid),
                                                          ptr(st.bwd_tiles(D)), n, H, D, cfg.neg_slope, cfg.p_attn, cfg.attn_seed,
                                                          cfg.attn_stream, ptr(alpha), ptr(alpha_d), ptr(elog), None, ldo, epi,
                                                          ptr(maskbits), None, ptr(o_hi), ptr(o_lo), ld16, ptr(bound), ptr(o_scale),
                                                          stream), "tx_gat_fused_fwd_staged")
                    else:
                        check(lib.tx_gat_fused_fwd_f16(ptr(ft), F_, ptr(al), ptr(ar), ptr(st.in_ptr), ptr(st.in_src), ptr(st.in_eid), n, H, D,
                                                       cfg.neg_slope, cfg.p_attn, cfg.attn_seed, cfg.attn_stream, ptr(alpha), ptr(alpha_d),
                                                       ptr(elog), ldo, epi, ptr(maskbits), ptr(o_hi), ptr(o_lo), ld16, ptr(bound),
                                                       ptr(o_scale), stream), "tx_gat_fused_fwd")
                    out16 = F16Pair(o_hi, o_lo, o_scale, F_ + pd)
                elif star and STAR_FWD_OUTPUT_LAYER and out_lo is None and n > 0:
                    sg = st.star
                    check(lib.tx_gat_star_fwd(ptr(ft), F_, ptr(al), ptr(ar), ptr(sg[0]), sg[1], sg[2], n, H, D, cfg.neg_slope, cfg.p_attn, cfg.attn_seed, cfg.attn_stream, ptr(alpha), ptr(alpha_d),
                                              ptr(elog), ptr(out), ldo, epi, ptr(maskbits), None, None, 0, None, None,
                                              ptr(_star_queue(dev)), stream), "tx_gat_star_fwd")
                elif STAGED_FWD and n > 0:
                    check(lib.tx_gat_fused_fwd_staged(ptr(ft), F_, ptr(al), ptr(ar), ptr(st.in_ptr), ptr(st.in_src), ptr(st.in_eid),
                                                      ptr(st.bwd_tiles(D)), n, H, D, cfg.neg_slope, cfg.p_attn, cfg.attn_seed,
                                                      cfg.attn_stream, ptr(alpha), ptr(alpha_d), ptr(elog), ptr(out), ldo, epi,
                                                      ptr(maskbits), ptr(out_lo), None, None, 0, None, None, stream),
                          "tx_gat_fused_fwd_staged")
                else:
                    check(lib.tx_gat_fused_fwd(ptr(ft), F_, ptr(al), ptr(ar), ptr(st.in_ptr), ptr(st.in_src), ptr(st.in_eid), n, H, D,
                                               cfg.neg_slope, cfg.p_attn, cfg.attn_seed, cfg.attn_stream, ptr(alpha), ptr(alpha_d),
                                               ptr(elog), ptr(out), ldo, epi, ptr(maskbits), ptr(out_lo), stream), "tx_gat_fused_fwd")
                if cfg.out_link is not None:
                    lk = cfg.out_link
                    lk.z_lo = out_lo
                    lk.z16 = out16
                    if maskbits is not None:
                        lk.mask, lk.heads, lk.dim, lk.act_slope, lk.p_drop, lk.applied = maskbits, H, D, cfg.act_slope, cfg.p_next, False
            else:
                a1 = torch.empty(n * H, **f32)
                a2 = torch.empty(n * H, **f32)
                check(lib.tx_gat_node_logits(ptr(ft), F_, ptr(al), ptr(ar), n, H, D, ptr(a1), ptr(a2), stream), "tx_gat_node_logits")
                check(lib.tx_gat_aggregate_fwd(ptr(ft), F_, ptr(al), ptr(ar), ptr(a1), ptr(a2), ptr(st.in_ptr), ptr(st.in_src),
                                               ptr(st.in_eid), n, st.e, H, D, cfg.neg_slope, cfg.p_attn, cfg.attn_seed,
                                               cfg.attn_stream, ptr(alpha), ptr(alpha_d), ptr(elog), ptr(out), ldo, epi, stream),
                      "tx_gat_aggregate_fwd")
        ctx.fused, ctx.maskbits = fused, maskbits
        ctx.st, ctx.cfg, ctx.pd = st, cfg, pd
        ctx.vocab = 0 if next_pos_table is None else int(next_pos_table.shape[0])
        ctx.save_for_backward(zsaved[0], zsaved[1], weight, al, ar, ft, alpha, alpha_d, elog,
                              out if (cfg.hidden and not fused) else None, pos32, zsaved[2], ft_amax, zsaved[3], zsaved[4], zsaved[5])
        ctx.attn_shape = attn_l.shape
        ctx.zshape = tuple(z.shape)
        return out

    @staticmethod
    def _native_saved(ctx):
        """The per-kernel backward's saved tensors as views of a native forward's workspace (rare fall-back: frozen weights, a
        consumer that did not apply the activation / dropout derivative)."""
        desc, state, ws, pd = ctx.native
        weight, attn_l, attn_r, tab, pos32, _ = ctx.saved_tensors
        cfg, st = ctx.cfg, ctx.st
        n = ctx.zshape[0]
        H, D, K = cfg.heads, cfg.dim, cfg.k
        F_ = H * D

        def view(src, p, count, dtype):
            nb = count * torch.empty((), dtype=dtype).element_size()
            off = p - src.data_ptr()
            return src[off:off + nb].view(dtype)
        if cfg.in_link is None:
            zsrc, ldz16 = ws, round8(K)
        else:
            zsrc, ldz16 = cfg.in_link.c_ws, int(state.ldz16)
        z_hi = view(zsrc, state.z_hi, n * ldz16, torch.float16).view(n, ldz16)
        z_lo = view(zsrc, state.z_lo, n * ldz16, torch.float16).view(n, ldz16)
        z_scale = view(zsrc, state.z_scale, 1, torch.float32)
        ldwt = int(state.ldwt)
        wt_hi = view(ws, state.wt_hi, K * ldwt, torch.float16).view(K, ldwt)
        wt_lo = view(ws, state.wt_lo, K * ldwt, torch.float16).view(K, ldwt)
        w_scale = view(ws, state.w_scale, 1, torch.float32)
        ft = view(ws, state.ft, n * F_, torch.float32).view(n, F_)
        ft_amax = view(ws, state.ft_amax, 1, torch.float32)
        alpha = view(ws, state.alpha, st.e * H, torch.float32)
        elog = view(ws, state.elog, st.e * H, torch.float32)
        alpha_d = view(ws, state.alpha_d, st.e * H, torch.float32)
        if state.maskbits:
            ctx.maskbits = view(ws, state.maskbits, int(_lib.load().tx_gat_fused_mask_words(n, H, D)), torch.int32)
        return (z_hi, z_lo, weight, attn_l.reshape(-1), attn_r.reshape(-1), ft, alpha, alpha_d, elog, None, pos32, z_scale, ft_amax,
                wt_hi, wt_lo, w_scale)

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        if ctx.native is not None:
            desc, state, ws, pd = ctx.native
            cfg, st = ctx.cfg, ctx.st
            lk_out = cfg.out_link
            pre = lk_out is not None and lk_out.applied
            if ctx.needs_input_grad[1] and (pre or not state.maskbits) and dout.dim() == 2:
                # ---- one native call: bounds / star backward / dW (+ attention rows) / d(attn) / d(z) (tx_layer.cu) ----
                weight, attn_l, attn_r, tab, pos32, z_in = ctx.saved_tensors
                n, ldz = ctx.zshape
                H, D, K = cfg.heads, cfg.dim, cfg.k
                F_ = H * D
                dev = dout.device
                f32 = dict(dtype=torch.float32, device=dev)
                with device_guard(dev):
                    Stats.sync_native_profiling()
                    dout = _rowmajor(dout)
                    if cfg.hidden and n > 1 and dout.stride(0) != round4(F_ + pd):
                        dout = dout.contiguous()  # the dropout counters of the forward epilogue are indexed with ITS row pitch
                    ldg = dout.stride(0) if n > 1 else dout.shape[1]
                    need_tab = cfg.hidden and ctx.needs_input_grad[4] and pd > 0
                    need_z = ctx.needs_input_grad[0]
                    g_amax = None
                    hand = st._dh_bound
                    st._dh_bound = None
                    if lk_out is not None and lk_out.c_bwd is not None and pre:
                        g_amax = lk_out.c_bwd[0]                            # max|d(z_next)|, measured by the layer above's d(z) GEMM
                    elif lk_out is not None and lk_out.dz_amax is not None and pre:
                        g_amax = lk_out.dz_amax.data_ptr()
                    elif not cfg.hidden and hand is not None and hand[0] == dout.data_ptr():
                        g_amax = hand[1].data_ptr()                         # published by the readout's backward
                    bws = torch.empty(int(lib.tx_gat_layer_bwd_bytes(ctypes.byref(desc))), dtype=torch.uint8, device=dev)
                    ldc = round4(K)
                    dw_ext = torch.empty((F_ + 2 * H, ldc), **f32)
                    # gradients go straight to their registered homes (a flat bucket's slices) when there are any, else to fresh
                    # CONTIGUOUS tensors - either way autograd adopts them as .grad without a copy
                    dw = grad_sink(weight, F_ * K)
                    dw = torch.empty((F_, K), **f32) if dw is None else dw.view(F_, K)
                    dal, dar = grad_sink(attn_l, F_), grad_sink(attn_r, F_)
                    if dal is None or dar is None:
                        both = torch.empty(2 * F_, **f32)
                        dal, dar = both[:F_], both[F_:]
                    dz = torch.empty((n, ldz), **f32) if need_z else None
                    dtab = None
                    if need_tab:
                        dtab = grad_sink(tab, ctx.vocab * pd)
                        dtab = torch.empty((ctx.vocab, pd), **f32) if dtab is None else dtab.view(ctx.vocab, pd)
                    dz_amax = ctypes.c_void_p()
                    prev = None if cfg.in_link is None else cfg.in_link.c_state
                    check(lib.tx_gat_layer_bwd(ctypes.byref(desc), ctypes.byref(state), None if prev is None else ctypes.byref(prev),
                                               ptr(dout), ldg, g_amax, ptr(bws), ptr(dz), ptr(dw_ext), ptr(dw), ptr(dal), ptr(dar), ptr(dtab),
                                               ctypes.byref(dz_amax), current_stream()), "tx_gat_layer_bwd")
                    if cfg.in_link is not None and need_z:
                        cfg.in_link.c_bwd = (dz_amax.value, bws)
                        cfg.in_link.applied = bool(prev is not None and prev.maskbits)
                        cfg.in_link.dz_amax = None
                return dz, dw, dal.view(ctx.attn_shape), dar.view(ctx.attn_shape), dtab, None, None, None
            saved = GatLayer._native_saved(ctx)
        else:
            saved = ctx.saved_tensors
        z0, z1, weight, al, ar, ft, alpha, alpha_d, elog, out, pos32, z_scale, ft_amax, wt_hi, wt_lo, wt_scale = saved
        st, cfg, pd = ctx.st, ctx.cfg, ctx.pd
        n, ldz = ctx.zshape
        H, D, K = cfg.heads, cfg.dim, cfg.k
        F_ = H * D
        dev = ft.device
        f32 = dict(dtype=torch.float32, device=dev)
        dtab = None
        with device_guard(dev):
            stream = current_stream()
            Stats.tag = cfg.tag
            dout = _rowmajor(dout)
            if cfg.hidden and n > 1 and dout.stride(0) != round4(F_ + pd):
                dout = dout.contiguous()      # the dropout counters of the forward epilogue are indexed with ITS row pitch
            ldg = dout.stride(0) if n > 1 else dout.shape[1]
            need_tab = cfg.hidden and ctx.needs_input_grad[4] and pd > 0
            nb = int(lib.tx_row_blocks(n))
            dft = torch.empty((n, F_), **f32)
            dft16 = None
            ds = torch.empty(st.e * H, **f32)
            da2 = torch.empty(n * H, **f32)
            hand = st._dh_bound               # the readout's hand-over is consumed (or dropped) by the FIRST layer backward after it
            st._dh_bound = None
            if ctx.fused:
                # g is read straight from d(z_next); dropout / leaky-relu derivative rebuilt from the forward's bit-planes
                if need_tab:
                    partial = torch.empty(nb * ctx.vocab * pd, **f32)
                    check(lib.tx_pos_grad_partials(ptr(dout), ldg, F_, ptr(pos32), n, pd, ctx.vocab, cfg.p_next, cfg.next_seed,
                                                   cfg.next_stream, ptr(partial), stream), "tx_pos_grad_partials")
                    dtab = _reduce_partials(lib, partial, nb, ctx.vocab * pd).view(ctx.vocab, pd)
                dft_lo = torch.empty_like(dft) if (FUSE_SPLIT and GEMM_BACKEND == "tf32x3") else None
                g_head_stride, g_scale = (D, 1.0) if cfg.hidden else (0, 1.0 / H)
                pre = cfg.out_link is not None and cfg.out_link.applied     # d(z_next) already carries the epilogue derivative
                if STAGED_BWD and (pre or ctx.maskbits is None):
                    # star backward (EgonetBatch) or the TMA-staged tile kernel (any batched graph)
                    f16out = FUSE_SPLIT and GEMM_BACKEND == "f16x3" and ft_amax is not None and n > 0
                    use_star = STAR_BWD and st.star_bwd is not None and dft_lo is None and n > 0 and H <= 64
                    need_w = ctx.needs_input_grad[1]
                    d_hi = d_lo = d_scale = bound = None
                    star_extra = 2 * H if (use_star and f16out) else 0
                    ld16 = round8(F_ + star_extra)
                    if f16out:
                        # dft goes out fp16-split.  Its scale needs an upper bound of |dft| BEFORE the kernel runs:
                        #   |sum_i alpha~_ij g_i| <= outdeg max|g| / (1 - p_attn),  |d alpha~| = |<g_i, ft_j>| <= D max|g| max|ft|,
                        #   |ds| <= 2 |d alpha~| / (1 - p_attn),  |da1_j| <= outdeg |ds|,  |da2_i| <= |ds|
                        # (rigorous, ~2^13 above the true maximum; the star backward tries DFT_OPTIMISM x max|g| first, see above)
                        if cfg.out_link is not None and cfg.out_link.dz_amax is not None and pre:
                            g_amax = cfg.out_link.dz_amax                 # published by the d(z) GEMM epilogue of the layer above
                        elif not cfg.hidden and hand is not None and hand[0] == dout.data_ptr():
                            g_amax = hand[1]                              # published by the readout's backward
                        else:
                            g_amax = absmax(dout, F_ if cfg.hidden else D)
                        deg = max(int(st.max_out_deg), 1)
                        slope = max(1.0, abs(cfg.neg_slope))
                        bound = torch.empty(4, **f32)
                        check(lib.tx_bound_dft(ptr(g_amax), ptr(ft_amax), ptr(al), ptr(ar), al.numel(), g_scale * deg / (1.0 - cfg.p_attn),
                                               g_scale * 2.0 * (deg + 1) * D * slope / (1.0 - cfg.p_attn),
                                               g_scale * DFT_OPTIMISM / (1.0 - cfg.p_attn) if use_star else 0.0, ptr(bound), stream),
                              "tx_bound_dft")
                        d_hi = torch.empty((n, ld16), dtype=torch.float16, device=dev)
                        d_lo = torch.empty((n, ld16), dtype=torch.float16, device=dev)
                        if ld16 > F_ and not use_star:
                            d_hi[:, F_:].zero_()
                            d_lo[:, F_:].zero_()
                        d_scale = torch.empty(1, **f32)
                        dft16 = F16Pair(d_hi, d_lo, d_scale, F_)
                    if use_star:
                        sg = st.star_bwd
                        partial = torch.empty(int(lib.tx_gat_star_bwd_partial_floats(sg[1], H, D)), **f32)
                        attn_from_gemm = f16out and need_w               # d(attn) from 2 H extra rows of the weight-gradient GEMM
                        da1 = None if attn_from_gemm else torch.empty(n * H, **f32)
                        if not attn_from_gemm:
                            star_extra = 0
                        flag = None if bound is None else bound[3:4].view(torch.int32)
                        check(lib.tx_gat_star_bwd(ptr(dout), ldg, g_head_stride, g_scale, ptr(ft), F_, ptr(alpha), ptr(alpha_d), ptr(elog),
                                                  ptr(al), ptr(ar), ptr(sg[0]), sg[1], sg[2], n, H, D, cfg.neg_slope, ptr(ds),
                                                  ptr(da1), None if attn_from_gemm else ptr(da2), None if f16out else ptr(dft), F_,
                                                  ptr(d_hi), ptr(d_lo), ld16, ptr(bound), ptr(flag), ptr(star_bwd_reruns(dev)) if f16out else None,
                                                  ptr(d_scale), ptr(partial), ptr(_star_counters(dev, sg[1] * H)),
                                                  ptr(_star_queue(dev)), stream), "tx_gat_star_bwd")
                        if attn_from_gemm:
                            dw, dz = _layer_gemms_bwd(dft, F_, (z0, z1, z_scale, wt_hi, wt_lo, wt_scale), weight.t(), K, ldz, cfg.dz_from, True,
                                                      ctx.needs_input_grad[0], cfg.in_link, None, dft16, extra_rows=star_extra)
                            v = dw[F_:]
                            dw = dw[:F_]
                            both = torch.empty(2 * F_, **f32)
                            check(lib.tx_attn_grad_from_v(ptr(weight), weight.stride(0), ptr(v), v.stride(0), H, D, K, ptr(bound[2:3]),
                                                          ptr(both), ptr(both[F_:]), stream), "tx_attn_grad_from_v")
                            return dz, dw, both[:F_].view(ctx.attn_shape), both[F_:].view(ctx.attn_shape), dtab, None, None, None
                        nbp = int(lib.tx_row_blocks(n))
                        partial = torch.empty(nbp * 2 * F_, **f32)
                        check(lib.tx_gat_attn_grad_partials(ptr(ft), F_, ptr(da1), ptr(da2), n, H, D, ptr(partial), stream),
                              "tx_gat_attn_grad_partials")
                        nbf = nbp
                    else:
                        tiles = st.bwd_tiles(D)
                        nbf = int(lib.tx_gat_fused_bwd_staged_blocks(n, H, D))
                        partial = torch.empty(nbf * 2 * F_, **f32)
                        check(lib.tx_gat_fused_bwd_staged(ptr(dout), ldg, g_head_stride, g_scale, ptr(ft), F_, ptr(alpha), ptr(alpha_d),
                                                          ptr(elog), ptr(al), ptr(ar), ptr(st.in_ptr), ptr(st.in_src), ptr(st.in_eid),
                                                          ptr(st.out_ptr), ptr(st.out_dst), ptr(st.out_slot), ptr(tiles), n, H, D,
                                                          cfg.neg_slope, cfg.p_attn, cfg.attn_seed, cfg.attn_stream, ptr(ds), ptr(da2),
                                                          ptr(dft), F_, ptr(dft_lo), ptr(d_hi), ptr(d_lo), ld16, ptr(bound), ptr(d_scale),
                                                          ptr(partial), stream), "tx_gat_fused_bwd_staged")
                else:
                    nbf = int(lib.tx_gat_fused_bwd_blocks(n, H))
                    partial = torch.empty(nbf * 2 * F_, **f32)
                    check(lib.tx_gat_fused_bwd(ptr(dout), ldg, g_head_stride, g_scale, None if pre else ptr(ctx.maskbits),
                                               1 if cfg.p_next > 0.0 else 0, cfg.act_slope, cfg.p_next if cfg.hidden else 0.0,
                                               ptr(ft), F_, ptr(alpha), ptr(alpha_d),
                                               ptr(elog), ptr(al), ptr(ar), ptr(st.in_ptr), ptr(st.in_src), ptr(st.in_eid),
                                               ptr(st.out_ptr), ptr(st.out_dst), ptr(st.out_slot), ptr(st.node_off), st.g, n, H, D,
                                               cfg.neg_slope, cfg.p_attn, cfg.attn_seed, cfg.attn_stream, ptr(ds), ptr(da2), ptr(dft),
                                               F_, ptr(dft_lo), ptr(partial), stream), "tx_gat_fused_bwd")
                both = _reduce_partials(lib, partial, nbf, 2 * F_)
                dal, dar = both[:F_].view(ctx.attn_shape), both[F_:].view(ctx.attn_shape)
            else:
                dft_lo = None
                if cfg.hidden:
                    if cfg.p_next > 0.0 or cfg.act_slope != 1.0 or need_tab:
                        if cfg.p_next > 0.0 or cfg.act_slope != 1.0:
                            dout = dout.clone()
                            ldg = dout.stride(0) if n > 1 else dout.shape[1]
                        partial = torch.empty(nb * ctx.vocab * pd, **f32) if need_tab else None
                        check(lib.tx_epilogue_bwd(ptr(dout), ldg, ptr(out), ptr(pos32) if pd > 0 else None, n, F_,
                                                  pd if need_tab else 0, ctx.vocab, cfg.act_slope, cfg.p_next, cfg.next_seed,
                                                  cfg.next_stream, ptr(partial), stream), "tx_epilogue_bwd")
                        if need_tab:
                            dtab = _reduce_partials(lib, partial, nb, ctx.vocab * pd).view(ctx.vocab, pd)
                    g, g_head_stride, g_scale = dout, D, 1.0
                else:
                    g, g_head_stride, g_scale = dout, 0, 1.0 / H
                check(lib.tx_gat_aggregate_bwd_dst(ptr(g), ldg, g_head_stride, g_scale, ptr(ft), F_, ptr(alpha), ptr(elog),
                                                   ptr(st.in_ptr), ptr(st.in_src), ptr(st.in_eid), n, H, D, cfg.neg_slope,
                                                   cfg.p_attn, cfg.attn_seed, cfg.attn_stream, ptr(ds), ptr(da2), stream),
                      "tx_gat_aggregate_bwd_dst")
                da1 = torch.empty(n * H, **f32)
                check(lib.tx_gat_aggregate_bwd_src(ptr(g), ldg, g_head_stride, g_scale, ptr(alpha_d), ptr(ds), ptr(da2), ptr(al),
                                                   ptr(ar), ptr(st.out_ptr), ptr(st.out_dst), ptr(st.out_slot), n, H, D, ptr(da1),
                                                   ptr(dft), F_, stream), "tx_gat_aggregate_bwd_src")
                dal = dar = None
                if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
                    partial = torch.empty(nb * 2 * F_, **f32)
                    check(lib.tx_gat_attn_grad_partials(ptr(ft), F_, ptr(da1), ptr(da2), n, H, D, ptr(partial), stream),
                          "tx_gat_attn_grad_partials")
                    both = _reduce_partials(lib, partial, nb, 2 * F_)
                    dal, dar = both[:F_].view(ctx.attn_shape), both[F_:].view(ctx.attn_shape)
            dw, dz = _layer_gemms_bwd(dft, F_, (z0, z1, z_scale, wt_hi, wt_lo, wt_scale), weight.t(), K, ldz, cfg.dz_from, ctx.needs_input_grad[1],
                                      ctx.needs_input_grad[0], cfg.in_link, dft_lo, dft16)
        return dz, dw, dal, dar, dtab, None, None, None


# --------------------------------------------------------------------------------------------------
# GCN layer: y = z W ; out = act(norm * sum_in(norm * y) + b) ; same epilogue
# --------------------------------------------------------------------------------------------------
@dataclass
class GcnLayerCfg:
    k: int
    dim: int
    hidden: bool = True
    act_slope: float = 1.0
    p_next: float = 0.0
    next_seed: int = 0
    next_stream: int = 0
    dz_from: int = 0
    tag: str = ""
    in_link: Optional[MaskLink] = None     # published by the previous layer (its epilogue produced this layer's z)
    out_link: Optional[MaskLink] = None    # published to the next layer


# GCN hidden layers on the fused epilogue the GAT path has (tx_gcn_aggregate_fwd_f16 / _bwd_f16): the output goes out as the next GEMM's
# fp16 pair + sign / keep bytes, the next layer's d(z) GEMM applies the activation / dropout derivative, d(y) goes out as a pair.
# TAXO_GCN_FUSED=0 -> fp32 round trips with a separate tx_epilogue_bwd pass (round 1)
GCN_FUSED = os.environ.get("TAXO_GCN_FUSED", "1") not in ("", "0")


class GcnLayer(Function):
    @staticmethod
    def forward(ctx, z, weight, bias, next_pos_table, st: GraphStructure, pos32, cfg: GcnLayerCfg):
        lib = _lib.load()
        _check_cuda(z, "z")
        n, ldz = z.shape
        D, K = cfg.dim, cfg.k
        dev = z.device
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = GEMM_BACKEND == "f16x3"
        ctx.native = None
        if (LAYER_CALL and GCN_FUSED and FUSE_SPLIT and FUSE_DZ_EPILOGUE and f16 and n > 0 and D % 4 == 0 and weight.shape[0] == K
                and weight.is_contiguous() and (bias is None or bias.is_contiguous()) and (not cfg.hidden or cfg.out_link is not None)
                and (next_pos_table is None or next_pos_table.is_contiguous())
                and (cfg.in_link.c_state is not None if cfg.in_link is not None
                     else (z.stride(1) == 1 and z.stride(0) % 4 == 0 and z.data_ptr() % 16 == 0))):
            # ---- one native call: split / weight split / GEMM / bound / aggregate + epilogue into one workspace (tx_layer.cu) ----
            with device_guard(dev):
                Stats.sync_native_profiling()
                norm = st.gcn_norm()
                pd = 0 if (next_pos_table is None or not cfg.hidden) else int(next_pos_table.shape[1])
                d = GcnLayerDesc()
                d.n, d.k, d.dim, d.pos_dim = n, K, D, pd
                d.vocab = 0 if pd == 0 else int(next_pos_table.shape[0])
                d.dz_from, d.max_in_deg, d.max_out_deg = cfg.dz_from, max(int(st.max_in_deg), 1), max(int(st.max_out_deg), 1)
                d.hidden, d.act_slope, d.p_next = (1 if cfg.hidden else 0), cfg.act_slope, (cfg.p_next if cfg.hidden else 0.0)
                d.next_seed, d.next_stream = cfg.next_seed, cfg.next_stream
                d.in_ptr, d.in_src, d.out_ptr, d.out_dst = st.in_ptr.data_ptr(), st.in_src.data_ptr(), st.out_ptr.data_ptr(), st.out_dst.data_ptr()
                d.pos = None if pd == 0 else pos32.data_ptr()
                d.norm, d.weight, d.ldw = norm.data_ptr(), weight.data_ptr(), weight.stride(0)
                d.bias = None if bias is None else bias.data_ptr()
                d.next_pos_table = None if pd == 0 else next_pos_table.data_ptr()
                d.tag = cfg.tag.encode()[:15]
                prev = None if cfg.in_link is None else cfg.in_link.c_state
                ws = torch.empty(int(lib.tx_gcn_layer_fwd_bytes(ctypes.byref(d), 1 if prev is None else 0)), dtype=torch.uint8, device=dev)
                state = GatLayerState()
                out = torch.empty((n, round4(D + pd)) if cfg.hidden else (n, D), **f32)     # hidden: a placeholder, never written
                check(lib.tx_gcn_layer_fwd(ctypes.byref(d), ptr(z) if prev is None else None, ldz, None if prev is None else ctypes.byref(prev),
                                           ptr(ws), ctypes.byref(state), None if cfg.hidden else ptr(out), current_stream()), "tx_gcn_layer_fwd")
            if cfg.out_link is not None:
                lk = cfg.out_link
                lk.c_state, lk.c_ws, lk.applied, lk.z16, lk.z_lo, lk.dz_amax, lk.c_bwd = state, ws, False, None, None, None, None
                lk.c_dims = (n, D + pd, round8(D + pd), int(lib.tx_gat_fused_mask_words(n, 1, D)))
                lk.mask = None
                if state.maskbits:
                    lk.heads, lk.dim, lk.act_slope, lk.p_drop = 1, D, cfg.act_slope, cfg.p_next
            ctx.native = (d, state, ws, pd)
            ctx.st, ctx.cfg, ctx.pd = st, cfg, pd
            ctx.vocab = 0 if next_pos_table is None else int(next_pos_table.shape[0])
            ctx.has_bias = bias is not None
            ctx.save_for_backward(weight, bias, next_pos_table, pos32, norm)
            ctx.zshape = tuple(z.shape)
            return out
        if cfg.in_link is not None:
            cfg.in_link.materialize()
        with device_guard(dev):
            stream = current_stream()
            Stats.tag = cfg.tag
            with timed_region("gemm_fwd"):
                y, zsaved, y_amax = _layer_gemms_fwd(z, K, weight.t(),         # torch.mm(h, W), model_zoo.py:37
                                                     z16=cfg.in_link.z16 if (cfg.in_link is not None and f16) else None,
                                                     w_rowmajor=weight if (f16 and weight.shape[0] == K) else None, w_is_nk=False)
            norm = st.gcn_norm()
            pd = 0 if next_pos_table is None else int(next_pos_table.shape[1])
            tab = None if next_pos_table is None else next_pos_table.contiguous()
            ldo = round4(D + pd) if cfg.hidden else D
            out = torch.empty((n, ldo), **f32)
            epi = GatEpilogue(mean_heads=0 if cfg.hidden else 1, act_slope=cfg.act_slope, next_pos_table=ptr(tab),
                              pos=ptr(pos32) if pd > 0 else None, pos_dim=pd, p_drop=cfg.p_next if cfg.hidden else 0.0,
                              seed=cfg.next_seed, stream_id=cfg.next_stream)
            b = None if bias is None else bias.contiguous()
            fused = (GCN_FUSED and FUSE_SPLIT and FUSE_DZ_EPILOGUE and f16 and cfg.hidden and cfg.out_link is not None and y_amax is not None
                     and n > 0 and D % 4 == 0 and y.stride(0) % 4 == 0)
            maskbits = None
            if fused:
                # the epilogue writes the next layer's input fp16-split; `out` itself is never written.  |norm_i sum_j norm_j y_j + b| <=
                # sqrt(max in-degree) max|y| + max|b|, appended rows <= max|P|, both / (1 - p_next)
                keep = 1.0 - cfg.p_next
                bound = torch.empty(1, **f32)
                check(lib.tx_bound_gcn(ptr(y_amax), (max(int(st.max_in_deg), 1) ** 0.5) / keep, ptr(b), 0 if b is None else b.numel(), 1.0 / keep,
                                       ptr(tab) if pd > 0 else None, tab.numel() if pd > 0 else 0, 1.0 / keep, ptr(bound), stream), "tx_bound_gcn")
                ld16 = round8(D + pd)
                o_hi = torch.empty((n, ld16), dtype=torch.float16, device=dev)
                o_lo = torch.empty((n, ld16), dtype=torch.float16, device=dev)
                o_scale = torch.empty(1, **f32)
                if cfg.act_slope != 1.0 or cfg.p_next > 0.0:
                    maskbits = torch.empty(int(lib.tx_gat_fused_mask_words(n, 1, D)), dtype=torch.int32, device=dev)
                check(lib.tx_gcn_aggregate_fwd_f16(ptr(y), y.stride(0), ptr(norm), ptr(b), ptr(st.in_ptr), ptr(st.in_src), n, D, ldo, epi,
                                                   ptr(o_hi), ptr(o_lo), ld16, ptr(bound), ptr(o_scale), ptr(maskbits), stream),
                      "tx_gcn_aggregate_fwd_f16")
                lk = cfg.out_link
                lk.z16, lk.z_lo, lk.c_state, lk.dz_amax, lk.applied = F16Pair(o_hi, o_lo, o_scale, D + pd), None, None, None, False
                lk.mask = maskbits
                if maskbits is not None:
                    lk.heads, lk.dim, lk.act_slope, lk.p_drop = 1, D, cfg.act_slope, cfg.p_next
            else:
                check(lib.tx_gcn_aggregate_fwd(ptr(y), y.stride(0), ptr(norm), ptr(b), ptr(st.in_ptr), ptr(st.in_src), n, D, ptr(out), ldo,
                                               epi, stream), "tx_gcn_aggregate_fwd")
        ctx.st, ctx.cfg, ctx.pd = st, cfg, pd
        ctx.vocab = 0 if next_pos_table is None else int(next_pos_table.shape[0])
        ctx.has_bias = bias is not None
        ctx.fused, ctx.has_mask = fused, maskbits is not None
        ctx.save_for_backward(zsaved[0], zsaved[1], weight, out if (cfg.hidden and not fused) else None, pos32, norm, zsaved[2], zsaved[3],
                              zsaved[4], zsaved[5])
        ctx.zshape = tuple(z.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        if ctx.native is not None:
            # ---- one native call: d(pos table) / d(bias) / bound / aggregate backward / dW / d(z) (tx_layer.cu) ----
            d, state, ws, pd = ctx.native
            weight, bias, tab, pos32, norm = ctx.saved_tensors
            st, cfg = ctx.st, ctx.cfg
            n, ldz = ctx.zshape
            D, K = cfg.dim, cfg.k
            dev = dout.device
            f32 = dict(dtype=torch.float32, device=dev)
            lk_out = cfg.out_link
            pre = lk_out is not None and lk_out.applied
            if state.maskbits and not pre:
                raise _lib.TaxoLibraryError("GcnLayer: the consumer of a fused GCN layer did not apply the activation / dropout derivative "
                                            "(set TAXO_LAYER_CALL=0 TAXO_GCN_FUSED=0 for this configuration)")
            with device_guard(dev):
                Stats.sync_native_profiling()
                dout = _rowmajor(dout)
                if cfg.hidden and n > 1 and dout.stride(0) != round4(D + pd):
                    dout = dout.contiguous()
                ldg = dout.stride(0) if n > 1 else dout.shape[1]
                need_tab = cfg.hidden and ctx.needs_input_grad[3] and pd > 0
                need_z = ctx.needs_input_grad[0]
                g_amax = None
                if lk_out is not None and lk_out.c_bwd is not None and pre:
                    g_amax = lk_out.c_bwd[0]
                elif lk_out is not None and lk_out.dz_amax is not None and pre:
                    g_amax = lk_out.dz_amax.data_ptr()
                bws = torch.empty(int(lib.tx_gcn_layer_bwd_bytes(ctypes.byref(d))), dtype=torch.uint8, device=dev)
                dwt = torch.empty((D, round4(K)), **f32)
                db = torch.empty(D, **f32) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
                dz = torch.empty((n, ldz), **f32) if need_z else None
                dtab = torch.empty((ctx.vocab, pd), **f32) if need_tab else None
                dz_amax = ctypes.c_void_p()
                prev = None if cfg.in_link is None else cfg.in_link.c_state
                check(lib.tx_gcn_layer_bwd(ctypes.byref(d), ctypes.byref(state), None if prev is None else ctypes.byref(prev), ptr(dout), ldg,
                                           g_amax, ptr(bws), ptr(dz), ptr(dwt), ptr(db), ptr(dtab), ctypes.byref(dz_amax), current_stream()),
                      "tx_gcn_layer_bwd")
                if cfg.in_link is not None and need_z:
                    cfg.in_link.c_bwd = (dz_amax.value, bws)
                    cfg.in_link.applied = bool(prev is not None and prev.maskbits)
                    cfg.in_link.dz_amax = None
            dw = dwt[:, :K].t() if ctx.needs_input_grad[1] else None            # weight is [K, D]
            return dz, dw, db, dtab, None, None, None
        z0, z1, weight, out, pos32, norm, z_scale, wt_hi, wt_lo, wt_scale = ctx.saved_tensors
        st, cfg, pd = ctx.st, ctx.cfg, ctx.pd
        n, ldz = ctx.zshape
        D, K = cfg.dim, cfg.k
        dev = norm.device
        f32 = dict(dtype=torch.float32, device=dev)
        dtab = db = None
        dy16 = None
        with device_guard(dev):
            stream = current_stream()
            Stats.tag = cfg.tag
            dout = _rowmajor(dout)
            if cfg.hidden and n > 1 and dout.stride(0) != round4(D + pd):
                dout = dout.contiguous()          # the dropout counters of the forward epilogue are indexed with ITS row pitch
            ldg = dout.stride(0) if n > 1 else dout.shape[1]
            pre = ctx.fused and (not ctx.has_mask or (cfg.out_link is not None and cfg.out_link.applied))
            if ctx.fused and not pre:
                raise _lib.TaxoLibraryError("GcnLayer: the consumer of a fused GCN layer did not apply the activation / dropout derivative "
                                            "(set TAXO_GCN_FUSED=0 for this configuration)")
            if cfg.hidden:
                need_tab = ctx.needs_input_grad[3] and pd > 0
                nb = int(lib.tx_row_blocks(n))
                if pre:
                    # d(z_next) already carries the epilogue derivative on the feature columns (the next layer's d(z) GEMM epilogue); the
                    # appended position rows get theirs here
                    if need_tab:
                        partial = torch.empty(nb * ctx.vocab * pd, **f32)
                        check(lib.tx_pos_grad_partials(ptr(dout), ldg, D, ptr(pos32), n, pd, ctx.vocab, cfg.p_next, cfg.next_seed,
                                                       cfg.next_stream, ptr(partial), stream), "tx_pos_grad_partials")
                        dtab = _reduce_partials(lib, partial, nb, ctx.vocab * pd).view(ctx.vocab, pd)
                elif cfg.p_next > 0.0 or cfg.act_slope != 1.0 or need_tab:
                    if cfg.p_next > 0.0 or cfg.act_slope != 1.0:
                        dout = dout.clone()
                        ldg = dout.stride(0) if n > 1 else dout.shape[1]
                    partial = torch.empty(nb * ctx.vocab * pd, **f32) if need_tab else None
                    check(lib.tx_epilogue_bwd(ptr(dout), ldg, ptr(out), ptr(pos32) if pd > 0 else None, n, D,
                                              pd if need_tab else 0, ctx.vocab, cfg.act_slope, cfg.p_next, cfg.next_seed,
                                              cfg.next_stream, ptr(partial), stream), "tx_epilogue_bwd")
                    if need_tab:
                        dtab = _reduce_partials(lib, partial, nb, ctx.vocab * pd).view(ctx.vocab, pd)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                nb = int(lib.tx_row_blocks(n))
                partial = torch.empty(nb * D, **f32)
                check(lib.tx_colsum_partials(ptr(dout), ldg, n, D, ptr(partial), stream), "tx_colsum_partials")
                db = _reduce_partials(lib, partial, nb, D)
            f16 = GEMM_BACKEND == "f16x3" and z0.shape[0] > 0
            dy = None
            if GCN_FUSED and FUSE_SPLIT and f16 and D % 4 == 0 and ldg % 4 == 0 and dout.data_ptr() % 16 == 0:
                # d(y) straight as the GEMMs' fp16 pair: |dy_j| = |norm_j sum_i norm_i g_i| <= out-degree max|g|
                if cfg.out_link is not None and cfg.out_link.dz_amax is not None and pre and cfg.hidden:
                    g_amax = cfg.out_link.dz_amax
                else:
                    g_amax = absmax(dout, D)
                bound = torch.empty(1, **f32)
                check(lib.tx_bound_gcn(ptr(g_amax), 0.5 * max(int(st.max_out_deg), 1), None, 0, 0.0, None, 0, 0.0, ptr(bound), stream), "tx_bound_gcn")
                ld16 = round8(D)
                d_hi = torch.empty((n, ld16), dtype=torch.float16, device=dev)
                d_lo = torch.empty((n, ld16), dtype=torch.float16, device=dev)
                d_scale = torch.empty(1, **f32)
                check(lib.tx_gcn_aggregate_bwd_f16(ptr(dout), ldg, ptr(norm), ptr(st.out_ptr), ptr(st.out_dst), n, D, ptr(d_hi), ptr(d_lo), ld16,
                                                   ptr(bound), ptr(d_scale), stream), "tx_gcn_aggregate_bwd_f16")
                dy16 = F16Pair(d_hi, d_lo, d_scale, D)
            else:
                dy = torch.empty((n, D), **f32)
                check(lib.tx_gcn_aggregate_bwd(ptr(dout), ldg, ptr(norm), ptr(st.out_ptr), ptr(st.out_dst), n, D, ptr(dy), D, stream),
                      "tx_gcn_aggregate_bwd")
            dwt, dz = _layer_gemms_bwd(dy, D, (z0, z1, z_scale, wt_hi, wt_lo, wt_scale), weight, K, ldz, cfg.dz_from, ctx.needs_input_grad[1],
                                       ctx.needs_input_grad[0], cfg.in_link, None, dy16)
            dw = None if dwt is None else dwt.t()                         # weight is [K, D]
        return dz, dw, db, dtab, None, None, None


# --------------------------------------------------------------------------------------------------
# y = a @ w  (the bilinear matching modules' projection, model_zoo.py:301-328), fp32-faithful on the tcgen05 GEMMs
# --------------------------------------------------------------------------------------------------
class DenseRight(Function):
    """y[m, r] = sum_l a[m, l] w[l, r] with its autograd GEMMs (da = dy w^T, dw = a^T dy) on the f16x3 kernels; the weight is
    split once per step in both orientations (tx_split_f16_weight)."""

    @staticmethod
    def forward(ctx, a, w):
        _check_cuda(a, "a")
        with device_guard(a.device):
            m, l = a.shape
            r = w.shape[1]
            ap = split_f16(a, l)
            wp, wt = split_f16_weight(w)                 # wp: [l, r] (input-gradient operand), wt: [r, l] (forward operand)
            y = gemm_nt_f16(ap, l, wt, r)
        ctx.save_for_backward(ap.hi, ap.lo, ap.scale, wp.hi, wp.lo, wp.scale)
        ctx.dims = (m, l, r)
        return y.contiguous() if y.shape[1] != r else y

    @staticmethod
    def backward(ctx, dy):
        a_hi, a_lo, a_sc, w_hi, w_lo, w_sc = ctx.saved_tensors
        m, l, r = ctx.dims
        da = dw = None
        with device_guard(dy.device):
            dp = split_f16(_rowmajor(dy), r)
            if ctx.needs_input_grad[0]:
                da = gemm_nt_f16(dp, r, F16Pair(w_hi, w_lo, w_sc, r), l)
            if ctx.needs_input_grad[1]:
                dw = gemm_tn_f16(F16Pair(a_hi, a_lo, a_sc, l), l, dp, r)
        return da, dw


def dense_right(a: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """a @ w through the library's fp32-faithful tensor-core GEMMs when they are the dense backend (else torch.mm)."""
    if GEMM_BACKEND == "f16x3" and a.is_cuda and a.shape[0] > 0 and a.dtype == torch.float32:
        return DenseRight.apply(a, w)
    return torch.mm(a, w)


# --------------------------------------------------------------------------------------------------
# Matching row-dot (+ exp) and InfoNCE (tx_match.cu)
# --------------------------------------------------------------------------------------------------
class MatchRowDot(Function):
    """scores[g, 0] = f(<u_g, q_g>), f = exp (LBM) or identity (BIM): the row-dot half of nn.Bilinear(l, r, 1) once u = e1 W[0]
    is known (reference model_zoo.py:301-328)."""

    @staticmethod
    def forward(ctx, u, q, apply_exp: bool):
        lib = _lib.load()
        _check_cuda(u, "projected graph embeddings")
        _check_cuda(q, "query features")
        u, q = _rowmajor(u), _rowmajor(q)
        G, r = u.shape
        if q.shape != u.shape:
            raise ValueError(f"match: {tuple(u.shape)} projected embeddings against {tuple(q.shape)} query features")
        scores = torch.empty((G, 1), dtype=torch.float32, device=u.device)
        with device_guard(u.device):
            check(lib.tx_match_rowdot_fwd(ptr(u), u.stride(0) if G > 1 else r, ptr(q), q.stride(0) if G > 1 else r, G, r,
                                          int(bool(apply_exp)), ptr(scores), current_stream()), "tx_match_rowdot_fwd")
        ctx.apply_exp = bool(apply_exp)
        ctx.save_for_backward(u, q, scores)
        return scores

    @staticmethod
    def backward(ctx, dscores):
        lib = _lib.load()
        u, q, scores = ctx.saved_tensors
        G, r = u.shape
        dscores = dscores.reshape(-1).contiguous()
        du = torch.empty((G, r), dtype=torch.float32, device=u.device) if ctx.needs_input_grad[0] else None
        dq = torch.empty((G, r), dtype=torch.float32, device=u.device) if ctx.needs_input_grad[1] else None
        with device_guard(u.device):
            check(lib.tx_match_rowdot_bwd(ptr(u), u.stride(0) if G > 1 else r, ptr(q), q.stride(0) if G > 1 else r, ptr(scores),
                                          ptr(dscores), G, r, int(ctx.apply_exp), ptr(du), r, ptr(dq), r, current_stream()),
                  "tx_match_rowdot_bwd")
        return du, dq, None


def match_rowdot(u: torch.Tensor, q: torch.Tensor, apply_exp: bool) -> torch.Tensor:
    return MatchRowDot.apply(u, q, apply_exp)


class InfoNCE(Function):
    """sum_q CE(output[q, :], target[q]) = F.cross_entropy(output, target, reduction="sum") (reference loss.py:52-57)."""

    @staticmethod
    def forward(ctx, output, target32):
        lib = _lib.load()
        _check_cuda(output, "scores")
        output = _rowmajor(output).contiguous()
        nq, m = output.shape
        if nq > 0 and m < 1:
            raise ValueError("info_nce_loss: no classes")
        work = torch.empty((2, max(nq, 1)), dtype=torch.float32, device=output.device)
        loss = torch.empty((), dtype=torch.float32, device=output.device)
        with device_guard(output.device):
            check(lib.tx_info_nce_fwd(ptr(output), nq, max(m, 1), ptr(target32), ptr(work[0]), ptr(work[1]), ptr(loss),
                                      current_stream()), "tx_info_nce_fwd")
        ctx.save_for_backward(output, work, target32)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        lib = _lib.load()
        output, work, target32 = ctx.saved_tensors
        nq, m = output.shape
        dloss = dloss.reshape(()).contiguous().to(torch.float32)
        dscores = torch.empty_like(output)
        with device_guard(output.device):
            check(lib.tx_info_nce_bwd(ptr(output), ptr(work[1]), nq, max(m, 1), ptr(target32), ptr(dloss), ptr(dscores),
                                      current_stream()), "tx_info_nce_bwd")
        return dscores, None


# --------------------------------------------------------------------------------------------------
# Readout
# --------------------------------------------------------------------------------------------------
class Readout(Function):
    @staticmethod
    def forward(ctx, h, pos_weight, st: GraphStructure, pos32, kind):
        lib = _lib.load()
        _check_cuda(h, "node states")
        h = _rowmajor(h)
        n, D = h.shape
        if n != st.n:
            raise ValueError(f"readout: {n} node rows for a graph with {st.n} nodes")
        width = 3 * D if kind == _lib.TX_READOUT_CONCAT else D
        hg = torch.empty((st.g, width), dtype=torch.float32, device=h.device)
        w = None if pos_weight is None else pos_weight.reshape(-1).contiguous()
        with device_guard(h.device):
            check(lib.tx_readout_fwd(kind, ptr(h), h.stride(0) if n > 1 else D, ptr(pos32), ptr(w), ptr(st.node_off), st.g, D,
                                     ptr(hg), width, current_stream()), "tx_readout_fwd")
        ctx.st, ctx.kind = st, kind
        st._dh_bound = None            # a bound left behind by an earlier backward on this structure must never be picked up
        ctx.wshape = None if pos_weight is None else pos_weight.shape
        ctx.save_for_backward(h, hg, w, pos32)
        return hg

    @staticmethod
    def backward(ctx, dhg):
        lib = _lib.load()
        h, hg, w, pos32 = ctx.saved_tensors
        st, kind = ctx.st, ctx.kind
        n, D = h.shape
        dhg = _rowmajor(dhg)
        dh = torch.empty((n, D), dtype=torch.float32, device=h.device)
        dw = None
        with device_guard(h.device):
            need_w = kind == _lib.TX_READOUT_WMEAN and ctx.needs_input_grad[1]
            nbr = int(lib.tx_readout_bwd_blocks(st.g))
            partial = torch.empty(nbr * 3, dtype=torch.float32, device=h.device) if need_w else None
            check(lib.tx_readout_bwd(kind, ptr(dhg), dhg.stride(0) if st.g > 1 else dhg.shape[1], ptr(h),
                                     h.stride(0) if n > 1 else D, ptr(hg), hg.shape[1], ptr(pos32), ptr(w), ptr(st.node_off),
                                     st.g, D, ptr(dh), D, ptr(partial), current_stream()), "tx_readout_bwd")
            if need_w:
                dw = _reduce_partials(lib, partial, nbr, 3).view(ctx.wshape)
            if GEMM_BACKEND == "f16x3" and st.g > 0:
                # every readout weight a_i / S, 1 / n_g is <= 1, so max|d(h)| <= max|d(hg)|: a 16 MB reduction instead of one over the
                # 74 MB d(h); the output layer's backward picks it up if it receives exactly this tensor
                st._dh_bound = (dh.data_ptr(), absmax(dhg))
        return dh, dw, None, None, None


# --------------------------------------------------------------------------------------------------
# Readout + bilinear matching as one native call per direction (tx_head_fwd / tx_head_bwd, tx_layer.cu)
# --------------------------------------------------------------------------------------------------
def head_native_ok(h: torch.Tensor, qf: torch.Tensor, kind: int, w: torch.Tensor) -> bool:
    """TaxoExpan.forward's readout + match (model.py:85-86) run as one call when the default dense back-end is on, the readout is a
    (weighted) mean and the matcher bilinear; everything else takes the two modules one after the other."""
    return (LAYER_CALL and GEMM_BACKEND == "f16x3" and kind in (_lib.TX_READOUT_MEAN, _lib.TX_READOUT_WMEAN) and h.is_cuda and qf.is_cuda
            and h.dtype == torch.float32 and qf.dtype == torch.float32 and h.dim() == 2 and qf.dim() == 2 and h.shape[0] > 0
            and qf.shape[0] > 0 and h.stride(1) == 1 and qf.stride(1) == 1 and w.dim() == 2 and w.is_contiguous()
            and w.shape[0] == h.shape[1] and w.shape[1] == qf.shape[1])


class ReadoutMatch(Function):
    """scores[g] = f(<readout(h)_g W, q_g>) (f = exp for LBM) with its whole autograd, one native call per direction; publishes
    hg through `holder['hg']` for callers that want the graph embeddings."""

    @staticmethod
    def forward(ctx, h, pos_weight, w, qf, st: GraphStructure, pos32, kind, apply_exp):
        lib = _lib.load()
        n, D = h.shape
        G, r = qf.shape
        if G != st.g or n != st.n:
            raise ValueError(f"readout + match: {n} node rows / {G} queries for a batch of {st.n} nodes / {st.g} graphs")
        dev = h.device
        with device_guard(dev):
            Stats.sync_native_profiling()
            d = HeadDesc()
            d.n, d.g, d.dim, d.r, d.kind, d.apply_exp = n, G, D, r, kind, 1 if apply_exp else 0
            pw = None if pos_weight is None else pos_weight.reshape(-1)
            d.pos = None if pos32 is None else pos32.data_ptr()
            d.node_off = st.node_off.data_ptr()
            d.pos_weight = None if pw is None else pw.data_ptr()
            d.w, d.ldw = w.data_ptr(), w.stride(0)
            d.tag = b"head"
            ws = torch.empty(int(lib.tx_head_fwd_bytes(ctypes.byref(d))), dtype=torch.uint8, device=dev)
            state = HeadState()
            scores = torch.empty((G, 1), dtype=torch.float32, device=dev)
            check(lib.tx_head_fwd(ctypes.byref(d), ptr(h), h.stride(0), ptr(qf), qf.stride(0), ptr(ws), ctypes.byref(state), ptr(scores),
                                  current_stream()), "tx_head_fwd")
        st._dh_bound = None
        ctx.native = (d, state, ws)
        ctx.st = st
        ctx.wshape = None if pos_weight is None else pos_weight.shape
        ctx.save_for_backward(h, qf, w, pos_weight, pos32)
        return scores

    @staticmethod
    def backward(ctx, dscores):
        lib = _lib.load()
        d, state, ws = ctx.native
        h, qf, w, pos_weight, pos32 = ctx.saved_tensors
        st = ctx.st
        n, D = h.shape
        r = qf.shape[1]
        dev = h.device
        f32 = dict(dtype=torch.float32, device=dev)
        with device_guard(dev):
            Stats.sync_native_profiling()
            dscores = dscores.reshape(-1).contiguous()
            bws = torch.empty(int(lib.tx_head_bwd_bytes(ctypes.byref(d))), dtype=torch.uint8, device=dev)
            dh = torch.empty((n, D), **f32)
            dw_buf = torch.empty((D, round4(r)), **f32)
            dw = grad_sink(w, D * r)
            dw = torch.empty((D, r), **f32) if dw is None else dw.view(D, r)
            need_pw = pos_weight is not None and ctx.needs_input_grad[1]
            dpw = None
            if need_pw:
                dpw = grad_sink(pos_weight, 3)
                dpw = torch.empty(3, **f32) if dpw is None else dpw
            amax = ctypes.c_void_p()
            check(lib.tx_head_bwd(ctypes.byref(d), ctypes.byref(state), ptr(h), h.stride(0), ptr(qf), qf.stride(0), ptr(dscores), ptr(bws),
                                  ptr(dh), ptr(dw_buf), ptr(dw), ptr(dpw), ctypes.byref(amax), current_stream()), "tx_head_bwd")
            # max|d(hg)| >= max|d(h)|: handed to the output layer's backward (it receives exactly this d(h))
            off = amax.value - bws.data_ptr()
            st._dh_bound = (dh.data_ptr(), bws[off:off + 4].view(torch.float32))
        return dh, (dpw.view(ctx.wshape) if need_pw else None), (dw if ctx.needs_input_grad[2] else None), None, None, None, None, None
