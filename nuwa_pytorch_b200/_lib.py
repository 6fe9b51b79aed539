"""ctypes binding of libnuwa_b200.so (the C-ABI declared in include/nuwa_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an exception is
raised.  PyTorch is used only for device memory and streams; raw device pointers cross the boundary.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnuwa_b200.so")

c_int, c_void_p, c_float, c_ll = ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_longlong
P = ctypes.POINTER


class LnParams(ctypes.Structure):  # mirrors nuwa_ln_params
    _fields_ = [("y", c_void_p), ("post_w", c_void_p), ("post_b", c_void_p), ("res_in", c_void_p),
                ("x_out", c_void_p), ("x_out_bf16", c_void_p), ("pre_w", c_void_p), ("pre_b", c_void_p),
                ("a_out", c_void_p), ("a_bs", c_ll), ("a_rs", c_int), ("a_t0", c_int), ("a_npos", c_int),
                ("shift", c_int), ("fmap", c_int), ("t0", c_int), ("B", c_int), ("nt", c_int), ("D", c_int),
                ("eps", c_float), ("t0_ptr", c_void_p), ("gather", c_int), ("shift_cache", c_void_p), ("sc_bs", c_ll)]


class AttnParams(ctypes.Structure):  # mirrors nuwa_attn_params
    _fields_ = [("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("o", c_void_p),
                ("q_bs", c_ll), ("k_bs", c_ll), ("v_bs", c_ll), ("o_bs", c_ll),
                ("q_rs", c_int), ("k_rs", c_int), ("v_rs", c_int), ("o_rs", c_int),
                ("B", c_int), ("nq", c_int), ("t0", c_int), ("H", c_int), ("dh", c_int), ("qscale", c_float),
                ("head_scale", c_void_p), ("talk", c_void_p), ("null_k", c_void_p), ("null_v", c_void_p),
                ("key_mask", c_void_p), ("mask_bs", c_int), ("bias", c_void_p), ("bias_nq", c_int), ("bias_nk", c_int),
                ("fmap", c_int), ("max_frames", c_int), ("nv", c_int), ("kt", c_int), ("kh", c_int), ("kw", c_int),
                ("dt", c_int), ("dh_", c_int), ("dw", c_int), ("causal", c_int), ("ck", c_int), ("cdil", c_int),
                ("jmax", c_int), ("nk_dense", c_int), ("t0_ptr", c_void_p)]


class EmbedParams(ctypes.Structure):  # mirrors nuwa_embed_params
    _fields_ = [("out", c_void_p), ("idx", c_void_p), ("idx_bs", c_ll), ("table", c_void_p), ("bos", c_void_p),
                ("ax1", c_void_p), ("ax2", c_void_p), ("ax3", c_void_p), ("d2", c_int), ("d3", c_int),
                ("has_bos", c_int), ("t0", c_int), ("B", c_int), ("nt", c_int), ("D", c_int), ("t0_ptr", c_void_p)]


class BgemmParams(ctypes.Structure):  # mirrors nuwa_bgemm_params
    _fields_ = [("A", c_void_p), ("B", c_void_p), ("C", c_void_p), ("M", c_int), ("N", c_int), ("K", c_int),
                ("a_trans", c_int), ("b_trans", c_int), ("lda", c_ll), ("ldb", c_ll), ("ldc", c_ll),
                ("batch1", c_int), ("batch2", c_int), ("a_s1", c_ll), ("a_s2", c_ll), ("b_s1", c_ll), ("b_s2", c_ll),
                ("c_s1", c_ll), ("c_s2", c_ll), ("alpha", c_float), ("c_bf16", c_int), ("accumulate", c_int)]


class LnBwdParams(ctypes.Structure):  # mirrors nuwa_lnbwd_params
    _fields_ = [("rows", c_int), ("nt", c_int), ("D", c_int), ("dout_f32", c_void_p), ("dout_bf16", c_void_p),
                ("unshift", c_int), ("fmap", c_int), ("x", c_void_p), ("x2", c_void_p), ("stable", c_int),
                ("w", c_void_p), ("eps", c_float), ("dx_bf16", c_void_p), ("dx_f32", c_void_p), ("dx2_f32", c_void_p),
                ("accumulate", c_int), ("dw", c_void_p), ("db", c_void_p), ("dcol", c_void_p)]


class EmbedBwdParams(ctypes.Structure):  # mirrors nuwa_embed_bwd_params
    _fields_ = [("dx", c_void_p), ("idx", c_void_p), ("idx_bs", c_ll), ("dtable", c_void_p), ("dbos", c_void_p),
                ("dax1", c_void_p), ("dax2", c_void_p), ("dax3", c_void_p), ("d2", c_int), ("d3", c_int),
                ("has_bos", c_int), ("B", c_int), ("nt", c_int), ("D", c_int), ("frac", c_float)]


class AttnRowsParams(ctypes.Structure):  # mirrors nuwa_attn_rows_params
    _fields_ = [("S", c_void_p), ("dPp", c_void_p), ("Pp", c_void_p), ("dS", c_void_p), ("talk", c_void_p),
                ("dtalk", c_void_p), ("B", c_int), ("H", c_int), ("nq", c_int), ("J", c_int), ("jp", c_int),
                ("out_scale", c_float), ("key_mask", c_void_p), ("mask_bs", c_int), ("has_null", c_int)]


class DecodeSub(ctypes.Structure):  # mirrors nuwa_decode_sub
    _fields_ = [("kind", c_int), ("shift", c_int), ("read", c_int), ("write", c_int),
                ("pre_w", c_void_p), ("pre_b", c_void_p), ("post_w", c_void_p), ("post_b", c_void_p),
                ("w_a", c_void_p), ("w_b", c_void_p), ("b_out", c_void_p), ("talk", c_void_p),
                ("null_k", c_void_p), ("null_v", c_void_p), ("cache", c_void_p), ("shift_cache", c_void_p),
                ("ip", c_int), ("kt", c_int), ("kh", c_int), ("kw", c_int), ("dt", c_int), ("dh_", c_int), ("dw", c_int),
                ("reserved", c_int)]


class DecodeParams(ctypes.Structure):  # mirrors nuwa_decode_params
    _fields_ = [("subs", c_void_p), ("nsubs", c_int),
                ("B", c_int), ("D", c_int), ("H", c_int), ("dh", c_int), ("npos", c_int), ("reversible", c_int),
                ("fmap", c_int), ("max_frames", c_int), ("causal", c_int), ("j3max", c_int),
                ("nk", c_int), ("key_mask", c_void_p), ("mask_bs", c_int),
                ("t_ptr", c_void_p), ("x_in", c_void_p), ("norm_w", c_void_p), ("norm_b", c_void_p),
                ("out_f32", c_void_p), ("out_bf16", c_void_p), ("w_logits", c_void_p), ("V", c_int),
                ("logits", c_void_p), ("y", c_void_p), ("act", c_void_p), ("actq", c_void_p), ("scores", c_void_p),
                ("barrier", c_void_p), ("kmax", c_int), ("jmax", c_int), ("split_small", c_int), ("split_ff", c_int), ("split_logits", c_int),
                ("max_ctas", c_int), ("prof", c_void_p), ("prof_cta", c_int), ("debug_flags", c_int)]


class OptChunk(ctypes.Structure):  # mirrors nuwa_opt_chunk
    _fields_ = [("offset", c_ll), ("len", c_int), ("weight_decay", c_int)]


class AdamWParams(ctypes.Structure):  # mirrors nuwa_adamw_params
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("chunks", c_void_p),
                ("nchunks", c_int), ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
                ("weight_decay", c_float), ("max_norm", c_float), ("grad_scale", c_float), ("sqnorm", c_void_p),
                ("step", c_int), ("step_ptr", c_void_p), ("zero_grad", c_int)]


# name -> argtypes (restype is int unless listed in _RESTYPES).  Must list EVERY symbol of include/nuwa_b200.h.
SIGNATURES = {
    "nuwa_abi_version": [],
    "nuwa_launch_count": [],
    "nuwa_strerror": [c_int],
    "nuwa_struct_sizes": [P(c_int)],
    "nuwa_gemm_prof_open": [],
    "nuwa_gemm_prof_attach": [c_void_p],
    "nuwa_gemm_prof_collect": [c_void_p, P(ctypes.c_double), P(c_float)],
    "nuwa_gemm_prof_bytes": [c_void_p],
    "nuwa_gemm_prof_close": [c_void_p],
    "nuwa_gemm_bf16": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                       c_void_p, c_int, c_int, c_int, c_void_p],
    "nuwa_conv2d_nhwc_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_int, c_int, c_void_p],
    "nuwa_sandwich_ln": [P(LnParams), c_void_p],
    "nuwa_stable_ln": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "nuwa_attn_sparse3dna": [P(AttnParams), c_void_p, c_void_p],
    "nuwa_attn_sparse3dna_halo": [P(AttnParams), c_void_p],
    "nuwa_attn_dense": [P(AttnParams), c_void_p, c_void_p],
    "nuwa_attn_sparse3dna_umma": [P(AttnParams), c_void_p],
    "nuwa_attn_cross2dna_umma": [P(AttnParams), c_void_p],
    "nuwa_attn_dense_pres": [P(AttnParams), c_void_p],
    "nuwa_attn_cross2dna": [P(AttnParams), c_void_p],
    "nuwa_embed_tokens": [P(EmbedParams), c_void_p],
    "nuwa_rotary_to_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_cross_entropy_mean": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "nuwa_sample_topk_gumbel": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                c_float, c_void_p],
    "nuwa_sample_topk_gumbel_at": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_float,
                                   c_float, c_void_p],
    "nuwa_cache_append": [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p],
    "nuwa_step_increment": [c_void_p, c_void_p],
    "nuwa_decode_stack": [P(DecodeParams), c_int, c_void_p],
    "nuwa_struct_sizes_decode": [P(c_int)],
    "nuwa_sqnorm_f32": [c_void_p, c_ll, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "nuwa_adamw_step": [P(AdamWParams), c_void_p],
    "nuwa_recon_loss_f32": [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "nuwa_struct_sizes_optim": [P(c_int)],
    "nuwa_nchw_f32_to_nhwc_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_nhwc_to_nchw_f32": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_im2col_nchw_f32": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_groupnorm_nhwc": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                            c_int, c_void_p],
    "nuwa_upsample2x_nhwc_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_vae_attn_prep": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "nuwa_vq_argmax": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_vq_argmax_tc_workspace": [c_int, c_int, c_int],
    "nuwa_vq_argmax_tc": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                          ctypes.c_ulonglong, c_void_p],
    "nuwa_split3_f32_bf16": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_linear_f32x3_workspace": [c_int, c_int],
    "nuwa_linear_f32x3": [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                          ctypes.c_ulonglong, c_void_p],
    "nuwa_gather_rows": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_conv1x1_nhwc_to_nchw": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    # ---- training (backward) ----
    "nuwa_gemm_bf16_splitk": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p],
    "nuwa_gemm_bf16_tn_splitk": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                 c_void_p],
    "nuwa_bgemm": [P(BgemmParams), c_void_p],
    "nuwa_ln_bwd_grid": [c_int],
    "nuwa_ln_bwd": [P(LnBwdParams), c_void_p],
    "nuwa_reduce_partials": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "nuwa_transpose_bf16": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p],
    "nuwa_geglu_fwd": [c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_geglu_bwd": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_ce_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "nuwa_embed_bwd": [P(EmbedBwdParams), c_void_p],
    "nuwa_rotary_bwd_to_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_add_rows_f32": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p],
    "nuwa_attn_bwd_rows": [P(AttnRowsParams), c_void_p],
    "nuwa_kv_full_build": [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                           c_int, c_void_p],
    "nuwa_kv_full_split": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int,
                           c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_mask_scores": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "nuwa_attn3dna_bwd_scores": [P(AttnParams), c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "nuwa_attn_dense_q1": [P(AttnParams), c_int, c_void_p],
    "nuwa_attn_dense_q1_bwd": [P(AttnParams), c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_void_p,
                               c_void_p, c_void_p],
    "nuwa_attn_dense_bwd_fused": [P(AttnParams), c_int, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p, c_float,
                                  c_void_p],
    "nuwa_attn3dna_bwd_dq_umma": [P(AttnParams), c_void_p, c_int, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_attnx2_bwd_dq_umma": [P(AttnParams), c_void_p, c_int, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_attn3dna_bwd_scores_umma": [P(AttnParams), c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "nuwa_attn3dna_bwd_dq": [P(AttnParams), c_void_p, c_int, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_attn3dna_bwd_dkdv": [P(AttnParams), c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll,
                               c_int, c_void_p],
    "nuwa_attnx2_bwd_scores_umma": [P(AttnParams), c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "nuwa_attnx2_bwd_scores": [P(AttnParams), c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "nuwa_attnx2_bwd_dq": [P(AttnParams), c_void_p, c_int, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_attnx2_bwd_dkdv": [P(AttnParams), c_int, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll,
                             c_int, c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "nuwa_attn_bwd_first_key": [c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_void_p, c_ll, c_void_p],
    "nuwa_attn3dna_bwd_first_key_finalize": [c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p],
    "nuwa_struct_sizes_bwd": [P(c_int)],
}
_RESTYPES = {"nuwa_strerror": ctypes.c_char_p, "nuwa_vq_argmax_tc_workspace": ctypes.c_ulonglong, "nuwa_linear_f32x3_workspace": ctypes.c_ulonglong, "nuwa_gemm_prof_bytes": ctypes.c_double, "nuwa_launch_count": ctypes.c_ulonglong, "nuwa_struct_sizes": None,
             "nuwa_gemm_prof_attach": None, "nuwa_gemm_prof_close": None, "nuwa_gemm_prof_open": c_void_p, "nuwa_struct_sizes_bwd": None, "nuwa_struct_sizes_decode": None, "nuwa_struct_sizes_optim": None}

_lib = None
# bumped by code that rewrites parameter storage behind autograd's back (optim.FusedAdamW): part of every packed-weight
# cache key next to (data_ptr, _version)
WEIGHTS_EPOCH = [0]


NUWA_ERR_INVALID = -1  # include/nuwa_b200.h


class NuwaB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Fails loudly when the extension is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NuwaB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nuwa_pytorch_b200 has no CPU / eager fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        sizes = (c_int * 3)()
        handle.nuwa_struct_sizes(sizes)
        mine = (ctypes.sizeof(LnParams), ctypes.sizeof(AttnParams), ctypes.sizeof(EmbedParams))
        if tuple(sizes) != mine:
            raise NuwaB200Error(f"struct layout mismatch between include/nuwa_b200.h {tuple(sizes)} and _lib.py {mine}")
        sizes = (c_int * 4)()
        handle.nuwa_struct_sizes_bwd(sizes)
        mine = tuple(ctypes.sizeof(c) for c in (BgemmParams, LnBwdParams, EmbedBwdParams, AttnRowsParams))
        if tuple(sizes) != mine:
            raise NuwaB200Error(f"backward struct layout mismatch: include/nuwa_b200.h {tuple(sizes)} vs _lib.py {mine}")
        sizes = (c_int * 2)()
        handle.nuwa_struct_sizes_decode(sizes)
        mine = (ctypes.sizeof(DecodeSub), ctypes.sizeof(DecodeParams))
        if tuple(sizes) != mine:
            raise NuwaB200Error(f"decode struct layout mismatch: include/nuwa_b200.h {tuple(sizes)} vs _lib.py {mine}")
        handle.nuwa_struct_sizes_optim(sizes)
        mine = (ctypes.sizeof(OptChunk), ctypes.sizeof(AdamWParams))
        if tuple(sizes) != mine:
            raise NuwaB200Error(f"optimizer struct layout mismatch: include/nuwa_b200.h {tuple(sizes)} vs _lib.py {mine}")
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().nuwa_strerror(code).decode()
        raise NuwaB200Error(f"{what} failed: {msg} (code {code})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NuwaB200Error("nuwa_pytorch_b200 kernels need CUDA tensors (there is no CPU path)")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().nuwa_launch_count())


class GemmProfiler:
    """Per-thread profiler of the tcgen05 GEMM / conv launches (context manager): events are recorded only for launches
    made by the thread that entered it.  collect() -> (launches, algorithmic flops, device ms), call after a sync;
    bytes() -> algorithmic HBM bytes (operands and outputs once each) of the launches the last collect() summed."""

    def __init__(self):
        self.h = lib().nuwa_gemm_prof_open()

    def __enter__(self):
        lib().nuwa_gemm_prof_attach(self.h)
        return self

    def __exit__(self, *a):
        lib().nuwa_gemm_prof_attach(None)

    def collect(self):
        fl, ms = ctypes.c_double(0.0), c_float(0.0)
        n = lib().nuwa_gemm_prof_collect(self.h, ctypes.byref(fl), ctypes.byref(ms))
        return int(n), float(fl.value), float(ms.value)

    def bytes(self):
        return float(lib().nuwa_gemm_prof_bytes(self.h))

    def close(self):
        if self.h:
            lib().nuwa_gemm_prof_close(self.h)
            self.h = None
