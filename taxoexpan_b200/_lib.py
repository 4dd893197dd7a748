"""ctypes binding of libtaxo_sm100.so (include/taxo_b200.h).

The product path has NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtaxo_sm100.so")

TX_READOUT_MEAN, TX_READOUT_WMEAN, TX_READOUT_CONCAT = 0, 1, 2


class TaxoLibraryError(RuntimeError):
    pass


class GatEpilogue(Structure):
    """struct tx_gat_epilogue"""
    _fields_ = [
        ("mean_heads", c_int32),
        ("act_slope", c_float),
        ("next_pos_table", c_void_p),
        ("pos", c_void_p),
        ("pos_dim", c_int64),
        ("p_drop", c_float),
        ("seed", c_uint64),
        ("stream_id", c_uint32),
    ]


P = c_void_p
I64 = c_int64
F32 = c_float


class GemmEpilogue(Structure):
    """struct tx_gemm_epilogue"""
    _fields_ = [("act_mask", c_void_p), ("heads", c_int64), ("dim", c_int64), ("mask_stride", c_int64), ("col0", c_int64),
                ("act_slope", c_float), ("p_drop", c_float), ("has_keep_plane", c_int32)]


class GatLayerDesc(Structure):
    """struct tx_gat_layer_desc"""
    _fields_ = [(k, c_int64) for k in ("n", "e", "k", "heads", "dim", "pos_dim", "vocab", "dz_from", "max_out_deg")] + [
        ("hidden", c_int32), ("neg_slope", c_float), ("p_attn", c_float), ("act_slope", c_float), ("p_next", c_float),
        ("dft_optimism", c_float), ("attn_seed", c_uint64), ("next_seed", c_uint64), ("attn_stream", c_uint32),
        ("next_stream", c_uint32), ("tasks_fwd", c_void_p), ("n_tasks_fwd", c_int64), ("chunk_fwd", c_int64),
        ("tasks_bwd", c_void_p), ("n_tasks_bwd", c_int64), ("chunk_bwd", c_int64), ("pos", c_void_p), ("queue", c_void_p),
        ("counters", c_void_p), ("reruns", c_void_p), ("weight", c_void_p), ("ldw", c_int64), ("attn_l", c_void_p),
        ("attn_r", c_void_p), ("next_pos_table", c_void_p), ("tag", ctypes.c_char * 16)]


class GcnLayerDesc(Structure):
    """struct tx_gcn_layer_desc"""
    _fields_ = [(k, c_int64) for k in ("n", "k", "dim", "pos_dim", "vocab", "dz_from", "max_in_deg", "max_out_deg")] + [
        ("hidden", c_int32), ("act_slope", c_float), ("p_next", c_float), ("next_seed", c_uint64), ("next_stream", c_uint32),
        ("in_ptr", c_void_p), ("in_src", c_void_p), ("out_ptr", c_void_p), ("out_dst", c_void_p), ("pos", c_void_p), ("norm", c_void_p),
        ("weight", c_void_p), ("ldw", c_int64), ("bias", c_void_p), ("next_pos_table", c_void_p), ("tag", ctypes.c_char * 16)]


class HeadDesc(Structure):
    """struct tx_head_desc"""
    _fields_ = [("n", c_int64), ("g", c_int64), ("dim", c_int64), ("r", c_int64), ("kind", c_int32), ("apply_exp", c_int32),
                ("pos", c_void_p), ("node_off", c_void_p), ("pos_weight", c_void_p), ("w", c_void_p), ("ldw", c_int64),
                ("tag", ctypes.c_char * 16)]


class HeadState(Structure):
    """struct tx_head_state"""
    _fields_ = [(k, c_void_p) for k in ("hg", "hg_hi", "hg_lo", "hg_scale", "w_hi", "w_lo", "w_scale", "u", "scores")]


class GatLayerState(Structure):
    """struct tx_gat_layer_state"""
    _fields_ = [("z_hi", c_void_p), ("z_lo", c_void_p), ("z_scale", c_void_p), ("ldz16", c_int64),
                ("wt_hi", c_void_p), ("wt_lo", c_void_p), ("w_scale", c_void_p), ("ldwt", c_int64),
                ("ft", c_void_p), ("ft_amax", c_void_p), ("alpha", c_void_p), ("alpha_d", c_void_p), ("elog", c_void_p),
                ("out_hi", c_void_p), ("out_lo", c_void_p), ("out_scale", c_void_p), ("ld16_out", c_int64), ("maskbits", c_void_p),
                ("heads", c_int64), ("dim", c_int64), ("act_slope", c_float), ("p_next", c_float)]


# name -> argtypes (all return int unless listed in _RESTYPES)
_SIGNATURES = {
    "tx_abi_version": [],
    "tx_last_error": [],
    "tx_target_arch": [],
    "tx_pdl_set": [ctypes.c_int],
    "tx_row_blocks": [I64],
    "tx_csr_workspace_bytes": [I64, I64, POINTER(c_int64)],
    "tx_build_csr_by_dst": [P, P, I64, I64, P, P, P, P, P, P],
    "tx_build_csr_by_src": [P, P, P, I64, I64, P, P, P, P, P],
    "tx_star_batch_structure": [P, P, P, P, I64, P, P, P, P, P, P, P, P, P, P],
    "tx_star_batch_plan": [P, P, I64, I64, I64, P, P, P, P, P],
    "tx_gather_rows": [P, I64, I64, P, I64, I64, P, I64, P],
    "tx_concat_pos_dropout_fwd": [P, I64, P, P, I64, I64, I64, P, I64, F32, c_uint64, c_uint32, P],
    "tx_concat_pos_dropout_f16": [P, I64, P, P, I64, I64, I64, I64, F32, c_uint64, c_uint32, P, P, P, I64, P, P],
    "tx_epilogue_bwd": [P, I64, P, P, I64, I64, I64, I64, F32, F32, c_uint64, c_uint32, P, P],
    "tx_reduce_partials": [P, I64, I64, P, P],
    "tx_reduce_partials_rows": [P, I64, I64, I64, I64, I64, P, I64, P],
    "tx_colsum_partials": [P, I64, I64, I64, P, P],
    "tx_gat_node_logits": [P, I64, P, P, I64, I64, I64, P, P, P],
    "tx_gat_aggregate_fwd": [P, I64, P, P, P, P, P, P, P, I64, I64, I64, I64, F32, F32, c_uint64, c_uint32, P, P, P, P,
                             I64, POINTER(GatEpilogue), P],
    "tx_gat_aggregate_bwd_dst": [P, I64, I64, F32, P, I64, P, P, P, P, P, I64, I64, I64, F32, F32, c_uint64, c_uint32,
                                 P, P, P],
    "tx_gat_aggregate_bwd_src": [P, I64, I64, F32, P, P, P, P, P, P, P, P, I64, I64, I64, P, P, I64, P],
    "tx_gat_attn_grad_partials": [P, I64, P, P, I64, I64, I64, P, P],
    "tx_gcn_norm": [P, I64, P, P],
    "tx_gcn_aggregate_fwd": [P, I64, P, P, P, P, I64, I64, P, I64, POINTER(GatEpilogue), P],
    "tx_gcn_aggregate_bwd": [P, I64, P, P, P, I64, I64, P, I64, P],
    "tx_readout_fwd": [c_int32, P, I64, P, P, P, I64, I64, P, I64, P],
    "tx_readout_bwd": [c_int32, P, I64, P, I64, P, I64, P, P, P, I64, I64, P, I64, P, P],
    "tx_dropout_keep_mask": [c_uint64, c_uint32, I64, I64, F32, P, P],
    "tx_readout_bwd_blocks": [I64],
    "tx_gat_fused_supported": [I64, I64, c_int32],
    "tx_gat_fused_mask_words": [I64, I64, I64],
    "tx_gat_fused_mask_ld": [I64, I64],
    "tx_gat_fused_bwd_blocks": [I64, I64],
    "tx_gat_fused_fwd": [P, I64, P, P, P, P, P, I64, I64, I64, F32, F32, c_uint64, c_uint32, P, P, P, P, I64,
                         POINTER(GatEpilogue), P, P, P],
    "tx_gat_fused_bwd": [P, I64, I64, F32, P, c_int32, F32, F32, P, I64, P, P, P, P, P, P, P, P, P, P, P, P, I64, I64, I64, I64,
                         F32, F32, c_uint64, c_uint32, P, P, P, I64, P, P, P],
    "tx_gat_bwd_tile_rows": [I64],
    "tx_gat_bwd_num_tiles": [I64, I64],
    "tx_gat_bwd_tiles": [P, I64, I64, P, P, I64, P, P],
    "tx_gat_fused_bwd_staged_blocks": [I64, I64, I64],
    "tx_gat_fused_bwd_staged": [P, I64, I64, F32, P, I64, P, P, P, P, P, P, P, P, P, P, P, P, I64, I64, I64, F32, F32, c_uint64, c_uint32,
                                P, P, P, I64, P, P, P, I64, P, P, P, P],
    "tx_gat_fused_fwd_staged": [P, I64, P, P, P, P, P, P, I64, I64, I64, F32, F32, c_uint64, c_uint32, P, P, P, P, I64,
                                POINTER(GatEpilogue), P, P, P, P, I64, P, P, P],
    "tx_gat_fused_fwd_f16": [P, I64, P, P, P, P, P, I64, I64, I64, F32, F32, c_uint64, c_uint32, P, P, P, I64,
                             POINTER(GatEpilogue), P, P, P, I64, P, P, P],
    "tx_pos_grad_partials": [P, I64, I64, P, I64, I64, I64, F32, c_uint64, c_uint32, P, P],
    "tx_split_tf32": [P, I64, I64, I64, P, P, I64, P],
    "tx_gemm_nt_tf32x3": [P, P, I64, P, P, I64, P, I64, I64, I64, I64, P],
    "tx_gemm_nt_tf32x3_ex": [P, P, I64, P, P, I64, P, I64, I64, I64, I64, POINTER(GemmEpilogue), P],
    "tx_gemm_tn_splits": [I64, I64, I64],
    "tx_gemm_tn_tf32x3": [P, P, I64, P, P, I64, P, I64, I64, I64, I64, I64, I64, P],
    "tx_absmax": [P, I64, I64, I64, P, P],
    "tx_bound_max2": [P, F32, P, I64, F32, P, P],
    "tx_bound_dft": [P, P, P, P, I64, F32, F32, F32, P, P],
    "tx_split_f16": [P, I64, I64, I64, P, P, P, I64, P, P],
    "tx_split_f16_weight": [P, I64, I64, I64, P, P, I64, P, P, I64, P, P, P],
    "tx_gemm_nt_f16x3": [P, P, I64, P, P, I64, P, P, P, I64, I64, I64, I64, POINTER(GemmEpilogue), P, P],
    "tx_gemm_tn_f16_splits": [I64, I64, I64],
    "tx_gemm_tn_f16x3": [P, P, I64, P, P, I64, P, P, P, I64, I64, I64, I64, I64, I64, P],
    "tx_gat_star_chunk": [],
    "tx_gat_star_max_chunks": [],
    "tx_gat_star_fwd": [P, I64, P, P, P, I64, I64, I64, I64, I64, F32, F32, c_uint64, c_uint32, P, P, P, P, I64,
                        POINTER(GatEpilogue), P, P, P, I64, P, P, P, P],
    "tx_gat_star_bwd_partial_floats": [I64, I64, I64],
    "tx_gat_star_bwd": [P, I64, I64, F32, P, I64, P, P, P, P, P, P, I64, I64, I64, I64, I64, F32, P, P, P, P, I64, P, P, I64,
                        P, P, P, P, P, P, P, P],
    "tx_attn_grad_from_v": [P, I64, P, I64, I64, I64, I64, P, P, P, P],
    "tx_gcn_aggregate_fwd_f16": [P, I64, P, P, P, P, I64, I64, I64, POINTER(GatEpilogue), P, P, I64, P, P, P, P],
    "tx_gcn_aggregate_bwd_f16": [P, I64, P, P, P, I64, I64, P, P, I64, P, P, P],
    "tx_bound_gcn": [P, F32, P, I64, F32, P, I64, F32, P, P],
    "tx_gat_layer_fwd_bytes": [POINTER(GatLayerDesc), c_int32],
    "tx_gat_layer_bwd_bytes": [POINTER(GatLayerDesc)],
    "tx_gat_layer_fwd": [POINTER(GatLayerDesc), P, I64, POINTER(GatLayerState), P, POINTER(GatLayerState), P, P],
    "tx_gat_layer_bwd": [POINTER(GatLayerDesc), POINTER(GatLayerState), POINTER(GatLayerState), P, I64, P, P, P, P, P, P, P, P,
                         POINTER(c_void_p), P],
    "tx_gcn_layer_fwd_bytes": [POINTER(GcnLayerDesc), c_int32],
    "tx_gcn_layer_bwd_bytes": [POINTER(GcnLayerDesc)],
    "tx_gcn_layer_fwd": [POINTER(GcnLayerDesc), P, I64, POINTER(GatLayerState), P, POINTER(GatLayerState), P, P],
    "tx_gcn_layer_bwd": [POINTER(GcnLayerDesc), POINTER(GatLayerState), POINTER(GatLayerState), P, I64, P, P, P, P, P, P,
                         POINTER(c_void_p), P],
    "tx_head_fwd_bytes": [POINTER(HeadDesc)],
    "tx_head_bwd_bytes": [POINTER(HeadDesc)],
    "tx_head_fwd": [POINTER(HeadDesc), P, I64, P, I64, P, POINTER(HeadState), P, P],
    "tx_head_bwd": [POINTER(HeadDesc), POINTER(HeadState), P, I64, P, I64, P, P, P, P, P, P, POINTER(c_void_p), P],
    "tx_set_after_star_bwd_event": [P],
    "tx_after_star_bwd_event_count": [],
    "tx_layer_launches": [c_int32],
    "tx_prof_enable": [c_int32],
    "tx_prof_clear": [],
    "tx_prof_count": [],
    "tx_prof_get": [I64, ctypes.c_char_p, ctypes.c_char_p, POINTER(c_float)],
    "tx_match_rowdot_fwd": [P, I64, P, I64, I64, I64, c_int32, P, P],
    "tx_match_rowdot_bwd": [P, I64, P, I64, P, P, I64, I64, c_int32, P, I64, P, I64, P],
    "tx_info_nce_fwd": [P, I64, I64, P, P, P, P, P],
    "tx_info_nce_bwd": [P, P, I64, I64, P, P, P, P],
}
_RESTYPES = {"tx_after_star_bwd_event_count": c_int64, "tx_last_error": c_char_p, "tx_target_arch": c_char_p, "tx_row_blocks": c_int64, "tx_readout_bwd_blocks": c_int64,
             "tx_gat_fused_mask_words": c_int64, "tx_gat_fused_mask_ld": c_int64, "tx_gat_fused_bwd_blocks": c_int64, "tx_gemm_tn_splits": c_int64,
             "tx_gat_bwd_tile_rows": c_int64, "tx_gat_bwd_num_tiles": c_int64, "tx_gat_fused_bwd_staged_blocks": c_int64,
             "tx_gemm_tn_f16_splits": c_int64, "tx_gat_star_chunk": c_int64, "tx_gat_star_max_chunks": c_int64,
             "tx_gat_star_bwd_partial_floats": c_int64, "tx_gat_layer_fwd_bytes": c_int64, "tx_gat_layer_bwd_bytes": c_int64, "tx_gcn_layer_fwd_bytes": c_int64, "tx_gcn_layer_bwd_bytes": c_int64, "tx_head_fwd_bytes": c_int64, "tx_head_bwd_bytes": c_int64,
             "tx_layer_launches": c_int64, "tx_prof_count": c_int64, "tx_prof_enable": None, "tx_prof_clear": None}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_raw = None

# names of ABI calls that do not enqueue GPU work
_NO_LAUNCH = {"tx_abi_version", "tx_last_error", "tx_target_arch", "tx_pdl_set", "tx_row_blocks", "tx_csr_workspace_bytes",
              "tx_readout_bwd_blocks", "tx_gat_fused_supported", "tx_gat_fused_mask_words", "tx_gat_fused_mask_ld", "tx_gat_fused_bwd_blocks", "tx_gemm_tn_splits",
              "tx_gat_bwd_tile_rows", "tx_gat_bwd_num_tiles", "tx_gat_fused_bwd_staged_blocks", "tx_gemm_tn_f16_splits", "tx_gat_star_chunk", "tx_gat_star_max_chunks", "tx_gat_star_bwd_partial_floats",
              # the per-layer calls enqueue several kernels each: they are counted through tx_layer_launches, timed through tx_prof_*
              "tx_gat_layer_fwd_bytes", "tx_gat_layer_bwd_bytes", "tx_gat_layer_fwd", "tx_gat_layer_bwd", "tx_layer_launches",
              "tx_gcn_layer_fwd_bytes", "tx_gcn_layer_bwd_bytes", "tx_gcn_layer_fwd", "tx_gcn_layer_bwd",
              "tx_head_fwd_bytes", "tx_head_bwd_bytes", "tx_head_fwd", "tx_head_bwd", "tx_set_after_star_bwd_event", "tx_after_star_bwd_event_count",
              "tx_prof_enable", "tx_prof_clear", "tx_prof_count", "tx_prof_get"}


class Stats:
    """Launch accounting + optional CUDA-event timing of every ABI call (used by bench.py; off by default)."""
    launches = 0          # ABI calls that enqueued kernels since the last reset
    profiling = False
    tag = ""
    events = []           # (name, tag, start_event, end_event)

    @classmethod
    def reset(cls):
        cls.launches = 0
        cls.events = []
        if _lib is not None:
            _lib.tx_layer_launches(1)
            _lib.tx_prof_clear()

    @classmethod
    def total_launches(cls):
        """ABI calls that enqueued kernels + the kernels enqueued inside the per-layer calls (tx_layer.cu) since the last reset."""
        return cls.launches + (int(_lib.tx_layer_launches(0)) if _lib is not None else 0)

    @classmethod
    def sync_native_profiling(cls):
        """Mirror `profiling` into the library (per-launch CUDA events inside the per-layer calls)."""
        if _lib is not None and cls._native_prof != cls.profiling:
            _lib.tx_prof_enable(1 if cls.profiling else 0)
            cls._native_prof = cls.profiling

    _native_prof = False

    @classmethod
    def timings_ms(cls):
        """{(name, tag): [ms, ...]} after a torch.cuda.synchronize()."""
        out = {}
        for name, tag, e0, e1 in cls.events:
            out.setdefault((name, tag), []).append(e0.elapsed_time(e1))
        if _lib is not None:
            nb, tb, ms = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64), c_float()
            for i in range(int(_lib.tx_prof_count())):
                check(_lib.tx_prof_get(i, nb, tb, ctypes.byref(ms)), "tx_prof_get")
                out.setdefault((nb.value.decode(), tb.value.decode()), []).append(float(ms.value))
        return out


class timed_region:
    """`with timed_region("mm", tag):` records CUDA events around non-ABI work (cuBLAS GEMMs) when profiling."""

    def __new__(cls, name, tag=None):
        if not Stats.profiling:
            return _NULL                      # nothing to record: no object, no event bookkeeping
        return super().__new__(cls)

    def __init__(self, name, tag=None):
        self.name, self.tag = name, tag

    def __enter__(self):
        if Stats.profiling:
            import torch
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if Stats.profiling:
            self.e1.record()
            Stats.events.append((self.name, Stats.tag if self.tag is None else self.tag, self.e0, self.e1))
        return False


class _Namespace:
    pass


def _wrap(fn, name):
    def call(*args):
        Stats.launches += 1
        if Stats.profiling:
            with timed_region(name):
                return fn(*args)
        return fn(*args)
    call.__name__ = name
    return call


def load():
    """Loads the shared library (once). Raises TaxoLibraryError with build instructions when it is absent."""
    global _lib, _raw
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TaxoLibraryError(
            f"{LIB_PATH} not found: build it with `python -m taxoexpan_b200.build` (needs nvcc). "
            "taxoexpan_b200 has no CPU or PyTorch fallback for the propagation/readout path.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise TaxoLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    ns = _Namespace()
    for name, argtypes in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise TaxoLibraryError(f"{LIB_PATH} does not export {name}; rebuild with `python -m taxoexpan_b200.build --force`") from e
        fn.argtypes = argtypes
        fn.restype = _RESTYPES[name] if name in _RESTYPES else ctypes.c_int
        setattr(ns, name, fn if name in _NO_LAUNCH else _wrap(fn, name))
    if lib.tx_abi_version() != 1:
        raise TaxoLibraryError(f"ABI version mismatch: library {lib.tx_abi_version()}, binding 1")
    _raw = lib
    _lib = ns
    return ns


def check(rc: int, what: str):
    if rc != 0:
        msg = load().tx_last_error().decode(errors="replace")
        raise TaxoLibraryError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None) as a plain int (ctypes converts it for the c_void_p parameters)."""
    return None if t is None else t.data_ptr()


class _NullContext:
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL = _NullContext()


def device_guard(device):
    """`with device_guard(t.device):` - torch.cuda.device(...) only when the tensor does not live on the current device (the context
    manager costs ~10 us and every autograd Function entered one or two of them per call)."""
    import torch
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(device)


_raw_stream = None


def current_stream():
    """cudaStream_t of torch's current stream on the current device (the raw-handle query: ~0.3 us instead of ~15 us for
    torch.cuda.current_stream().cuda_stream, which was a quarter of the host time of a training step)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return c_void_p(_raw_stream(torch.cuda.current_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)
