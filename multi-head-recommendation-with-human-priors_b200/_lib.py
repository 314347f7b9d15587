"""ctypes binding of libb200rec.so (the C ABI declared in include/b200rec.h).

There is no CPU fallback: a missing library or a failing kernel raises.  torch is used only to
own device memory and to name the current CUDA stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200rec.so")

F32, BF16 = 0, 1
(EPI_STORE, EPI_ACCUM, EPI_SILU_DUAL, EPI_BIAS_RESID, EPI_RESBLOCK, EPI_GT_BITS, EPI_FOLD_HEADS, EPI_NCE_EXP,
 EPI_FOLD_ITEMS) = range(9)

_lib = None
launches = 0  # number of C-ABI kernel entry points invoked (bench.py's gpu_launches claim)
gemm_timing = None  # bench.py sets this to a list: (start_event, end_event, flops) per GEMM launch
gemm_timing_external = False  # True while capturing an instrumented CUDA graph: events become external record nodes


def _timing_events():
    if gemm_timing_external:
        return (torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


class B200RecError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_major", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_major", C.c_int),
        ("in_dtype", C.c_int),
        ("C", C.c_void_p), ("ldc", C.c_int64), ("c_dtype", C.c_int),
        ("C2", C.c_void_p), ("ldc2", C.c_int64), ("c2_dtype", C.c_int),
        ("epilogue", C.c_int),
        ("alpha", C.c_float),
        ("alpha_dev", C.c_void_p),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("n_split", C.c_int), ("c_split_stride", C.c_int64), ("c2_split_stride", C.c_int64),
        ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_size_t),
        ("fold_hp", C.c_int),
        ("fold_head_on", C.c_void_p), ("fold_head_cat", C.c_void_p), ("fold_item_tags", C.c_void_p),
        ("fold_id_offset", C.c_int64), ("fold_id_stride", C.c_int64),
        ("fold_thr", C.c_void_p), ("fold_cnt", C.c_void_p), ("fold_keys", C.c_void_p), ("fold_cap", C.c_int),
        ("fold_groups", C.c_int),
        ("gt_row", C.c_void_p), ("gt_col", C.c_void_p),
        ("row_scale", C.c_void_p),
        ("nce_mref", C.c_void_p), ("nce_thr", C.c_void_p), ("nce_stats", C.c_void_p), ("nce_logit_scale", C.c_void_p),
        ("fold_on_bits", C.c_void_p),
    ]


class NcePosRefJob(C.Structure):
    _fields_ = [("q_hat", C.c_void_p), ("p_mask", C.c_uint32), ("tok_ok_col", C.c_int), ("pos_cos", C.c_void_p),
                ("mref", C.c_void_p), ("thr", C.c_void_p)]


class NceCombineJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("stats", "E", "same_bits", "row_any", "pos_cos", "mref", "q_hat", "coef", "loss",
                                          "g0", "dscale", "rank0", "nvalid", "row_scale", "qs")]


class NcePosBwdJob(C.Structure):
    _fields_ = [("g0", C.c_void_p), ("q_hat", C.c_void_p), ("d_qhat", C.c_void_p)]


_P, _I, _L, _F, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_SIGS = {
    "b200rec_version": (C.c_int, []),
    "b200rec_device_is_sm100": (C.c_int, []),
    "b200rec_gather_rows": (C.c_int, [_P, _I, _P, _L, _P, _I, _P]),
    "b200rec_embed_tokens": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "b200rec_gather_l2norm": (C.c_int, [_P, _P, _I, _P, _L, _P, _I, _P, _P]),
    "b200rec_l2norm_bwd": (C.c_int, [_P, _I, _P, _P, _L, _I, _P, _I, _P]),
    "b200rec_pos_emb_grad": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "b200rec_resblock_bwd": (C.c_int, [_P, _P, _I, _L, _I, _I, _P, _P, _P]),
    "b200rec_scatter_add_workspace_bytes": (_Z, [_L]),
    "b200rec_scatter_add_sorted": (C.c_int, [_P, _L, _P, _I, _P, _P, _P, _P, _Z, _P]),
    "b200rec_scatter_add_sorted_peer": (C.c_int, [_P, _L, _I, _I, _P, _I, _P, _P, _P, _P, _Z, _P]),
    "b200rec_gather_rows_sharded": (C.c_int, [_P, _I, _I, _P, _L, _P, _P]),
    "b200rec_ipc_export": (C.c_int, [_P, _P, _P]),
    "b200rec_ipc_import": (C.c_int, [_P, _P]),
    "b200rec_ipc_close": (C.c_int, [_P]),
    "b200rec_rows_to_dense": (C.c_int, [_P, _P, _P, _L, _I, _P, _I, _P]),
    "b200rec_layernorm_fwd": (C.c_int, [_P, _I, _I, _F, _P, _I, _P, _P, _P]),
    "b200rec_layernorm_bwd": (C.c_int, [_P, _I, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "b200rec_gate_ln_fwd": (C.c_int, [_P, _I, _P, _I, _I, _F, _P, _I, _P, _P, _F, C.c_uint32, C.c_uint32, _P, _P]),
    "b200rec_gate_ln_bwd": (C.c_int, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _P, _P, _I, _F, C.c_uint32, C.c_uint32, _P,
                                      _P]),
    "b200rec_counter_add": (C.c_int, [_P, _L, _P]),
    "b200rec_cast": (C.c_int, [_P, _L, _P, _I, _P]),
    "b200rec_colsum_workspace_bytes": (_Z, [_I, _I]),
    "b200rec_colsum": (C.c_int, [_P, _I, _I, _I, _I, _P, _I, _P, _Z, _P]),
    "b200rec_reduce_sum": (C.c_int, [_P, _L, _F, _P, _I, _P]),
    "b200rec_gemm": (C.c_int, [C.POINTER(GemmArgs), _P]),
    "b200rec_gemm_grouped": (C.c_int, [C.POINTER(GemmArgs), _I, _P]),
    "b200rec_gemm_force_bn": (None, [_I]),
    "b200rec_gemm_force_ctas": (None, [_I]),
    "b200rec_gemm_use_pdl": (None, [_I]),
    "b200rec_gemm_use_tail_split": (None, [_I]),
    "b200rec_hstu_attn_fwd": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P]),
    "b200rec_hstu_attn_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _I, _P,
                                        _P, _P, _P, _P]),
    "b200rec_hstu_attn_bias_fwd": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _P]),
    "b200rec_hstu_attn_bias_ws_floats": (_Z, [_I, _I, _I]),
    "b200rec_hstu_attn_bias_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P,
                                             _P, _P, _P, _P, _P]),
    "b200rec_hstu_attn_tc_fwd": (C.c_int, [_P, _I, _P, _P, _I, _I, _I, _I, _F, _P, _P]),
    "b200rec_hstu_attn_tc_bwd": (C.c_int, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P]),
    "b200rec_hstu_attn_seq_fwd": (C.c_int, [_P, _I, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P]),
    "b200rec_hstu_attn_seq_bwd": (C.c_int, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _P]),
    "b200rec_nce_loss_fwd": (C.c_int, [_P, _L, _I, _P, _P, _P, _P, _L, _P, _I, _I, _P, _P, _I, _I, _I, C.c_uint32,
                                       _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P]),
    "b200rec_nce_ihn_loss_fwd": (C.c_int, [_P, _L, _I, _P, _P, _L, _P, _I, _I, _P, _P, _I, _I, _I, C.c_uint32,
                                           _P, _I, _I, _P, _P, _F, _P, _P, _P, _P, _P, _P, _L, _P]),
    "b200rec_switch_rows": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "b200rec_switch_loss": (C.c_int, [_P, _L, _I, _I, _I, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _F, _F, _P,
                                      _P, _P, _P, _P]),
    "b200rec_switch_bwd": (C.c_int, [_P, _P, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "b200rec_switch_pos_grad": (C.c_int, [_P, _I, _I, _I, _I, _P, _P]),
    "b200rec_gemm_nce_parts": (C.c_int, [_I]),
    "b200rec_tail_norm": (C.c_int, [_P, _L, _I, _I, _P, _P]),
    "b200rec_prefix_aug": (C.c_int, [_P, _L, _I, _I, _P, _P]),
    "b200rec_gt_bits_verify": (C.c_int, [_P, _L, _I, _I, _P, _P, _I, _F, _P, _P]),
    "b200rec_gt_bits_verify_sets": (C.c_int, [_P, _L, _I, _I, _P, _P, _I, _F, _P, _I, _L, _L, _L, _P]),
    "b200rec_nce_pos_ref": (C.c_int, [_P, _L, _P, _I, _P, _P, _I, _I, _I, C.c_uint32, _P, _I, _I, _P, _P, _P, _P, _P]),
    "b200rec_nce_combine": (C.c_int, [_P, _I, _P, _L, _I, _P, _P, _P, _P, _P, _L, _I, _P, _P, _I, _I, _I, _P, _P, _P, _P,
                                      _P, _P, _P, _P, _P, _L, _P]),
    "b200rec_nce_pos_ref_grouped": (C.c_int, [_P, _I, _L, _P, _I, _P, _P, _I, _I, _I, _P, _I, _P, _P]),
    "b200rec_nce_combine_grouped": (C.c_int, [_P, _I, _I, _L, _I, _L, _I, _P, _P, _I, _I, _I, _L, _P, _L, _P]),
    "b200rec_nce_pos_bwd_q_grouped": (C.c_int, [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _P, _P, _L, _P]),
    "b200rec_nce_pos_bwd_t_grouped": (C.c_int, [_P, _I, _L, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "b200rec_nce_coefs": (C.c_int, [_P, _P, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P]),
    "b200rec_nce_topk_logs": (C.c_int, [_P, _P, _I, _I, _P, _P, _P]),
    "b200rec_nce_count": (C.c_int, [_P, _P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "b200rec_nce_coef": (C.c_int, [_P, _P, _F, _I, _P, _P]),
    "b200rec_nce_pos_bwd_q": (C.c_int, [_P, _P, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P, _L, _P]),
    "b200rec_nce_pos_bwd_t": (C.c_int, [_P, _P, _L, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "b200rec_topk_workspace_bytes": (_Z, [_I, _L]),
    "b200rec_score_mask_topk": (C.c_int, [_P, _L, _I, _I, _L, _I, _P, _P, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _Z,
                                          _P]),
    "b200rec_adamw_tick_hist": (C.c_int, [_P, _P, _I, _F, _F, _F, _P]),
    "b200rec_adamw_rows_catchup": (C.c_int, [_P, _P, _P, _L, _I, _P, _L, _P, _P, _I, _P, _F, _F, _F, _F, _L, _L, _P]),
    "b200rec_adamw_rows_lazy": (C.c_int, [_P, _P, _P, _L, _I, _P, _P, _P, _L, _P, _P, _I, _P, _F, _F, _F, _F, _F, _P]),
    "b200rec_build_train_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _I, _I, _P, _P, _I, _F, _P, _I,
                                            C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P]),
    "b200rec_build_eval_batch": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "b200rec_topk_from_candidates": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _L, _L, _P, _P, _P, _P, _P]),
    "b200rec_topk_select": (C.c_int, [_P, _P, _I, _L, _L, _I, _P, _P, _L, _L, _P, _P, _P, _P]),
    "b200rec_apply_score_masks": (C.c_int, [_P, _L, _I, _I, _L, _P, _P, _P, _P]),
    "b200rec_hit_matrix": (C.c_int, [_P, _P, _I, _I, _I, _P, _I, _P, _P]),
    "b200rec_adamw": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P, _P]),
    "b200rec_adamw_tick": (C.c_int, [_P, _F, _F, _P]),
    "b200rec_adamw_multi": (C.c_int, [_P, _P, _I, _F, _F, _F, _F, _F, _I, _F, _P, _P]),
    "b200rec_group_pairs_count": (C.c_int, [_P, _L, _I, _P, _P]),
    "b200rec_group_pairs_emit": (C.c_int, [_P, _P, _P, _L, _I, _P, _P]),
    "b200rec_louvain_best_move": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _L, _L, C.c_double, _I, _P, _P]),
    "b200rec_louvain_apply": (C.c_int, [_P, _P, _P, _P, _P, _L, _P, _P]),
    "b200rec_comi_tanh": (C.c_int, [_P, _L, _P]),
    "b200rec_comi_tanh_bwd": (C.c_int, [_P, _P, _L, _P]),
    "b200rec_comi_pool_fwd": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "b200rec_comi_select_fwd": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "b200rec_comi_select_bwd": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "b200rec_comi_pool_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "b200rec_comi_rr": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "b200rec_adamw_rows": (C.c_int, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _F, _F, _F, _F, _F, _I, _F, _P, _P]),
}
EXPORTS = ["b200rec_last_error"] + sorted(_SIGS)


def lib():
    """Loads the shared library once; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise B200RecError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C multi-head-recommendation-with-human-priors_b200/csrc). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        l.b200rec_last_error.restype = C.c_char_p
        l.b200rec_last_error.argtypes = []
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(rc, name):
    if rc != 0:
        raise B200RecError(f"{name} failed ({rc}): {lib().b200rec_last_error().decode()}")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dt(t_or_dtype):
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise B200RecError(f"unsupported dtype {d}")


def call(name, *args):
    global launches
    launches += 1
    _check(getattr(lib(), name)(*args), name)


_TAIL_WS = {}
TAIL_WS_BYTES = 48 << 20


def _tail_workspace(device):
    """fp32 scratch for the GEMM tail split, one per device: the GEMMs of a process are issued in order (one compute
    stream, or a graph replayed on it), so consecutive launches can share it; allocated on first use, which the
    graphed steps reach in their eager warm-up, before any capture."""
    key = device.index
    ws = _TAIL_WS.get(key)
    if ws is None:
        ws = _TAIL_WS[key] = torch.empty(TAIL_WS_BYTES // 4, dtype=torch.float32, device=device)
    return ws


def gemm(A, B, C_out, M, N, K, *, lda, ldb, ldc, a_major=0, b_major=0, epilogue=EPI_STORE, alpha=1.0,
         alpha_dev=None, bias=None, resid=None, ldr=0, C2=None, ldc2=0, n_split=0, c_dtype=None, fold=None,
         splitk_ws=None, gt=None, fold_items=None):
    """C[M,N] = epi(A[M,K] @ B[N,K]^T).  A/B are tensors (or views) whose data_ptr is element (0,0)."""
    global launches
    a = GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_major = A.data_ptr(), lda, a_major
    a.B, a.ldb, a.b_major = B.data_ptr(), ldb, b_major
    a.in_dtype = dt(A)
    if dt(B) != a.in_dtype:
        raise B200RecError("gemm: A and B dtypes differ")
    a.C, a.ldc = ptr(C_out), ldc
    raw_out = epilogue in (EPI_GT_BITS, EPI_FOLD_HEADS, EPI_FOLD_ITEMS)
    a.c_dtype = F32 if raw_out else (dt(C_out) if c_dtype is None else c_dtype)
    a.C2, a.ldc2 = ptr(C2), ldc2
    a.c2_dtype = dt(C2) if (C2 is not None and not raw_out) else F32
    if splitk_ws is not None and ldc == N:
        a.splitk_ws, a.splitk_ws_bytes = splitk_ws.data_ptr(), splitk_ws.numel() * splitk_ws.element_size()
    elif a.in_dtype == BF16 and epilogue <= EPI_RESBLOCK and M > 128:
        ws = _tail_workspace(A.device)     # tail split of the last partial wave (include/b200rec.h)
        a.splitk_ws, a.splitk_ws_bytes = ws.data_ptr(), ws.numel() * 4
    if fold is not None:   # (hp, head_on u8[M], head_cat i32[hp] or None, item_tags u32[N] or None, id_offset, id_stride)
        a.fold_hp, a.fold_head_on, a.fold_head_cat, a.fold_item_tags = fold[0], ptr(fold[1]), ptr(fold[2]), ptr(fold[3])
        a.fold_id_offset, a.fold_id_stride = fold[4], fold[5]
        if len(fold) > 6:      # streamed variant: (thr f32[users], cnt u32[users], keys u64[users, cap], cap)
            a.fold_thr, a.fold_cnt, a.fold_keys, a.fold_cap = ptr(fold[6]), ptr(fold[7]), ptr(fold[8]), fold[9]
            a.fold_groups = fold[10] if len(fold) > 10 else 1
    if fold_items is not None:   # (hp, on_bits u32[users], head_cat i32[hp] or None, item_tags u32[M] or None, id_offset,
        f = fold_items           #  id_stride[, thr f32[users], cnt u32[users], keys u64[users, cap], cap])
        a.fold_hp, a.fold_on_bits, a.fold_head_cat, a.fold_item_tags = f[0], ptr(f[1]), ptr(f[2]), ptr(f[3])
        a.fold_id_offset, a.fold_id_stride = f[4], f[5]
        if len(f) > 6:
            a.fold_thr, a.fold_cnt, a.fold_keys, a.fold_cap = ptr(f[6]), ptr(f[7]), ptr(f[8]), f[9]
    if gt is not None:     # GT_BITS upper-bound variant: (row vector fp32[M], column vector fp32[N])
        a.gt_row, a.gt_col = gt[0].data_ptr(), gt[1].data_ptr()
    a.epilogue, a.alpha = epilogue, alpha
    a.alpha_dev = ptr(alpha_dev)
    a.bias, a.resid, a.ldr = ptr(bias), ptr(resid), ldr
    a.n_split, a.c_split_stride, a.c2_split_stride = n_split, 0, 0
    launches += 1
    if gemm_timing is not None:
        e0, e1 = _timing_events()
        e0.record()
        _check(lib().b200rec_gemm(C.byref(a), stream()), "b200rec_gemm")
        e1.record()
        gemm_timing.append((e0, e1, 2.0 * M * N * K, (M, N, K, 1, epilogue, a_major, b_major)))
        return
    _check(lib().b200rec_gemm(C.byref(a), stream()), "b200rec_gemm")


def gemm_grouped(problems, M, N, K, *, lda, ldb, ldc, a_major=0, b_major=0, epilogue=EPI_STORE, alpha=1.0,
                 alpha_dev=None, row_scales=None, nce=None, nce_logit_scale=None):
    """Same-shape problems [(A, B, C), ...] with non-aliasing outputs in ONE persistent launch (16 per launch):
    C_g[M,N] = epi(A_g[M,K] @ B_g[N,K]^T), epilogue STORE or ACCUM (optional per-row factors `row_scales[g]`), or
    NCE_EXP with nce[g] = (mref, thr, stats) and the logit_scale device scalar."""
    global launches
    n = len(problems)
    if n == 0:
        return
    arr = (GemmArgs * n)()
    for a, (A, B, C_out) in zip(arr, problems):
        a.M, a.N, a.K = M, N, K
        a.A, a.lda, a.a_major = A.data_ptr(), lda, a_major
        a.B, a.ldb, a.b_major = B.data_ptr(), ldb, b_major
        a.in_dtype = dt(A)
        if dt(B) != a.in_dtype:
            raise B200RecError("gemm: A and B dtypes differ")
        a.C, a.ldc, a.c_dtype = C_out.data_ptr(), ldc, (F32 if epilogue == EPI_GT_BITS else dt(C_out))
        a.c2_dtype = F32
        a.epilogue, a.alpha, a.alpha_dev = epilogue, alpha, ptr(alpha_dev)
    if row_scales is not None:
        for a, rs in zip(arr, row_scales):
            a.row_scale = ptr(rs)
    if nce is not None:
        for a, (mref, thr, stats) in zip(arr, nce):
            a.nce_mref, a.nce_thr, a.nce_stats = mref.data_ptr(), thr.data_ptr(), stats.data_ptr()
            a.nce_logit_scale = nce_logit_scale.data_ptr()
    launches += (n + 15) // 16 if arr[0].in_dtype == BF16 else n
    if gemm_timing is not None:
        e0, e1 = _timing_events()
        e0.record()
        _check(lib().b200rec_gemm_grouped(arr, n, stream()), "b200rec_gemm_grouped")
        e1.record()
        gemm_timing.append((e0, e1, 2.0 * M * N * K * n, (M, N, K, n, epilogue, a_major, b_major)))
        return
    _check(lib().b200rec_gemm_grouped(arr, n, stream()), "b200rec_gemm_grouped")


def job_array(cls, rows):
    """ctypes array of job structs from dicts (tensors -> device pointers, None -> NULL)."""
    arr = (cls * len(rows))()
    for a, r in zip(arr, rows):
        for k, v in r.items():
            setattr(a, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    return arr


def call_grouped(name, arr, *args):
    """One C-ABI call over a job array; counts ceil(n / 16) launches (gpu_launches claim of bench.py)."""
    global launches
    launches += (len(arr) + 15) // 16
    _check(getattr(lib(), name)(C.cast(arr, C.c_void_p), len(arr), *args), name)


_COLSUM_WS = {}
_COLSUM_WS_RETIRED = []          # outgrown workspaces: a captured graph may still hold their address, so never freed
COLSUM_WS_BYTES = 32 << 20


def colsum(x, rows, cols, ldx, out, accumulate=False):
    """out[j] (+)= sum_i x[i, j]; deterministic, one launch.  The workspace is persistent per device: its tile counters
    start at zero and every call leaves them at zero (calls of one stream run in order).  Its address must stay valid for
    every CUDA graph that captured a call, so it is allocated once at a size no training shape of this repo outgrows
    (32 MiB = 256-row slabs x columns x 4 B); if a call still needs more, the old buffer is retired, not freed: a graph
    captured earlier keeps replaying into it (it is self-contained: counters zero on entry and on exit)."""
    nbytes = lib().b200rec_colsum_workspace_bytes(rows, cols)
    ws = _COLSUM_WS.get(x.device.index)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _COLSUM_WS_RETIRED.append(ws)
        ws = _COLSUM_WS[x.device.index] = torch.zeros(max(2 * nbytes, COLSUM_WS_BYTES), dtype=torch.uint8, device=x.device)
    call("b200rec_colsum", x.data_ptr(), dt(x), ldx, rows, cols, out.data_ptr(), 1 if accumulate else 0,
         ws.data_ptr(), ws.numel(), stream())
