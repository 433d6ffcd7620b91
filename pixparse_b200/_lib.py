"""ctypes binding of the C-ABI library (include/pixparse_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PIXPARSE_B200_LIB") or os.path.join(_HERE, "csrc", "libpixparse_b200.so")

_lib = None

EPI_STORE_BF16 = 0
EPI_GELU_BF16 = 1
EPI_RESID_F32 = 2
EPI_DGELU_BF16 = 3
EPI_REDUCE_F32 = 4
EPI_STORE_F32 = 5
ABI_VERSION = 3      # B200_ABI_VERSION of include/pixparse_b200.h


class B200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle. Raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} not found: build it with `python -m pixparse_b200.build` "
                "(there is no CPU / PyTorch fallback for the Cruller hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
        # A/B switches for bring-up (scripts/gpu_ab.sh): kernel selection only, results are equivalent
        if os.environ.get("PIXPARSE_B200_ATTN_BWD_QUERY_MAJOR"):
            _lib.b200_debug_attention_bwd_query_major(int(os.environ["PIXPARSE_B200_ATTN_BWD_QUERY_MAJOR"]))
        if os.environ.get("PIXPARSE_B200_GEMM_SINGLE_CTA"):
            _lib.b200_debug_gemm_single_cta(int(os.environ["PIXPARSE_B200_GEMM_SINGLE_CTA"]))
    return _lib


_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_longlong
_F = ctypes.c_float
_U = ctypes.c_uint


class _Args(ctypes.Structure):
    """Base of the POD argument structs: field order and C types mirror include/pixparse_b200.h exactly
    (tests/test_abi.py parses the header and compares); `struct_size` is filled in here and checked by the library."""

    def __init__(self, **kw):
        super().__init__()
        names = {f[0] for f in self._fields_}
        for k, v in kw.items():
            if k not in names:
                raise TypeError(f"{type(self).__name__} has no field {k!r}")
            setattr(self, k, v)
        self.struct_size = ctypes.sizeof(self)


class GemmArgs(_Args):
    _fields_ = [("struct_size", _U), ("epilogue", _I), ("a", _P), ("lda", _L), ("a_mn_major", _I), ("b", _P), ("ldb", _L),
                ("b_mn_major", _I), ("m", _I), ("n", _I), ("k", _I), ("out", _P), ("ldo", _L), ("out2", _P), ("ldo2", _L),
                ("bias", _P), ("aux", _P), ("ld_aux", _L), ("bias_grad", _P), ("splits", _I), ("block_n", _I),
                ("drop_p", _F), ("drop_seed", _U), ("tile_counter", _P), ("tail_workspace", _P), ("tail_workspace_bytes", _L)]


class AttentionFwdArgs(_Args):
    _fields_ = [("struct_size", _U), ("batch", _I), ("q", _P), ("ldq", _L), ("q_bstride", _L), ("q_col0", _I),
                ("k_col0", _I), ("k", _P), ("ldk", _L), ("k_bstride", _L), ("v", _P), ("ldv", _L), ("v_bstride", _L),
                ("v_col0", _I), ("heads", _I), ("out", _P), ("ld_out", _L), ("out_bstride", _L), ("lse", _P),
                ("key_mask", _P), ("key_mask_bstride", _L), ("sq", _I), ("sk", _I), ("head_dim", _I), ("causal", _I),
                ("scale", _F), ("drop_p", _F), ("drop_seed", _U), ("reserved", _I)]


class AttentionBwdArgs(_Args):
    _fields_ = [("struct_size", _U), ("batch", _I), ("q", _P), ("ldq", _L), ("k", _P), ("ldk", _L), ("v", _P), ("ldv", _L),
                ("q_col0", _I), ("k_col0", _I), ("v_col0", _I), ("do_col0", _I), ("o", _P), ("ld_o", _L), ("d_o", _P),
                ("ld_do", _L), ("lse", _P), ("dq", _P), ("ld_dq", _L), ("dk", _P), ("ld_dk", _L), ("dv", _P),
                ("ld_dv", _L), ("dq_col0", _I), ("dk_col0", _I), ("dv_col0", _I), ("heads", _I), ("workspace", _P),
                ("sq", _I), ("sk", _I), ("head_dim", _I), ("causal", _I), ("scale", _F), ("drop_p", _F),
                ("drop_seed", _U), ("reserved", _I)]


class LayerNormFwdArgs(_Args):
    _fields_ = [("struct_size", _U), ("rows", _I), ("x", _P), ("gamma", _P), ("beta", _P), ("y_bf16", _P), ("y_f32", _P),
                ("mean", _P), ("rstd", _P), ("dim", _I), ("eps", _F), ("drop_p", _F), ("drop_seed", _U)]


class LayerNormBwdArgs(_Args):
    _fields_ = [("struct_size", _U), ("rows", _I), ("dy_bf16", _P), ("dy_f32", _P), ("dres_f32", _P), ("x", _P),
                ("mean", _P), ("rstd", _P), ("gamma", _P), ("dx_f32", _P), ("dx_bf16", _P), ("dgamma", _P),
                ("dbeta", _P), ("dim", _I), ("in_p", _F), ("in_seed", _U), ("out_p", _F), ("out_seed", _U),
                ("reserved", _I)]


class AdamWArgs(_Args):
    _fields_ = [("struct_size", _U), ("num_segments", _I), ("params", _P), ("grads", _P), ("exp_avg", _P),
                ("exp_avg_sq", _P), ("params_bf16", _P), ("n", _L), ("segments", _P), ("norm_stats", _P),
                ("grad_scale", _F), ("lr", _F), ("beta1", _F), ("beta2", _F), ("eps", _F), ("step", _I),
                ("zero_grad", _I), ("reserved", _I)]


class DecodeLinearArgs(_Args):
    _fields_ = [("struct_size", _U), ("m", _I), ("x", _P), ("ldx", _L), ("w", _P), ("ldw", _L), ("bias", _P), ("resid", _P),
                ("ld_resid", _L), ("out_bf16", _P), ("out_f32", _P), ("ldo", _L), ("pos", _P), ("out_pos_stride", _L),
                ("argmax_partial", _P), ("n", _I), ("k", _I), ("act", _I), ("n_split", _I), ("out2_bf16", _P), ("ldo2", _L)]


class DecodeAttentionArgs(_Args):
    _fields_ = [("struct_size", _U), ("batch", _I), ("q", _P), ("ldq", _L), ("k", _P), ("v", _P), ("ld_kv", _L),
                ("kv_bstride", _L), ("out", _P), ("ld_out", _L), ("pos", _P), ("key_ids", _P), ("ld_ids", _L),
                ("pad_id", _L), ("q_col0", _I), ("k_col0", _I), ("v_col0", _I), ("heads", _I), ("head_dim", _I),
                ("sk", _I), ("scale", _F), ("reserved", _I)]


# C struct name (header) -> ctypes mirror
STRUCTS = {"B200GemmArgs": GemmArgs, "B200AttentionFwdArgs": AttentionFwdArgs, "B200AttentionBwdArgs": AttentionBwdArgs,
           "B200LayerNormFwdArgs": LayerNormFwdArgs, "B200LayerNormBwdArgs": LayerNormBwdArgs, "B200AdamWArgs": AdamWArgs,
           "B200DecodeLinearArgs": DecodeLinearArgs, "B200DecodeAttentionArgs": DecodeAttentionArgs}

# name -> argtypes; mirrors include/pixparse_b200.h (tests/test_abi.py compares every prototype and struct field)
SIGNATURES = {
    "b200_abi_version": [],
    "b200_device_check": [],
    "b200_debug_gemm_desc": [_I, _I, _I, _I, _I, _I],
    "b200_debug_gemm_single_cta": [_I],
    "b200_debug_gemm_tail_split": [_I],
    "b200_debug_attention_bwd_query_major": [_I],
    "b200_gemm_bf16": [ctypes.POINTER(GemmArgs), _P],
    "b200_attention_fwd": [ctypes.POINTER(AttentionFwdArgs), _P],
    "b200_attention_bwd": [ctypes.POINTER(AttentionBwdArgs), _P],
    "b200_layernorm_fwd": [ctypes.POINTER(LayerNormFwdArgs), _P],
    "b200_layernorm_bwd": [ctypes.POINTER(LayerNormBwdArgs), _P],
    "b200_adamw_step": [ctypes.POINTER(AdamWArgs), _P],
    "b200_colsum_bf16": [_P, _L, _I, _I, _P, _P],
    "b200_patch_unfold": [_P, _P, _I, _I, _I, _I, _I, _L, _P],
    "b200_tokens_assemble": [_P, _P, _P, _P, _I, _I, _I, _P],
    "b200_tokens_assemble_bwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "b200_embed_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "b200_embed_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _L, _P],
    "b200_set_pdl": [_I],
    "b200_cast_f32_bf16": [_P, _P, _L, _P],
    "b200_reduce_shards": [_P, _P, _L, _I, _I, _F, _L, _P],
    "b200_ce_prepare": [_P, _I, _L, _P, _P],
    "b200_ce_fwd_bwd": [_P, _L, _P, _P, _L, _P, _P, _I, _I, _L, _F, _P],
    "b200_grad_norm": [_P, _L, _P, _P, _F, _F, _P],
    "b200_grad_norm_workspace_floats": [],
    "b200_preprocess_pages": [_P, _I, _I, _I, _L, _P, _I, _I, _F, _F, _P, _P],
    "b200_decode_linear": [ctypes.POINTER(DecodeLinearArgs), _P],
    "b200_decode_linear_ctas": [_I],
    "b200_decode_attention": [ctypes.POINTER(DecodeAttentionArgs), _P],
    "b200_decode_embed": [_P, _L, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "b200_decode_finalize": [_P, _I, _P, _L, _P, _P, _I, _L, _P],
}
# entry points whose return type is not int
RESTYPES = {"b200_last_error": ctypes.c_char_p, "b200_attention_bwd_workspace_bytes": ctypes.c_longlong,
            "b200_preprocess_workspace_bytes": ctypes.c_longlong, "b200_gemm_tail_workspace_bytes": ctypes.c_longlong}
OTHER_SIGNATURES = {"b200_last_error": [], "b200_attention_bwd_workspace_bytes": [_I, _I, _I],
                    "b200_preprocess_workspace_bytes": [_I, _I], "b200_gemm_tail_workspace_bytes": []}


def _declare(l):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    for name, argtypes in OTHER_SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = RESTYPES[name]
    got = l.b200_abi_version()
    if got != ABI_VERSION:
        raise B200Error(f"{LIB_PATH} implements C-ABI version {got}, this package binds version {ABI_VERSION}: rebuild it "
                        "with `python -m pixparse_b200.build`")


def check(rc, what):
    if rc != 0:
        msg = lib().b200_last_error().decode("utf-8", "replace")
        raise B200Error(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# kernels launched per C-ABI call (grad_norm: partial + final; attention_bwd: prep + main + dq convert)
_KERNELS_PER_CALL = {"b200_grad_norm": 2, "b200_attention_bwd": 3}
_launches = 0
_profile_shapes = bool(int(os.environ.get("PIXPARSE_B200_PROFILE_SHAPES", "0")))      # per-(M, N, K) GEMM rows in profiles
_profile = None     # {name: [(start_event, end_event, flops), ...]} while profile_ops() is active


def reset_launch_count():
    global _launches
    _launches = 0


def launch_count():
    return _launches


def add_launches(n):
    """Kernel launches that did not go through call(): replays of a captured CUDA graph (n = kernel nodes x replays;
    negative to take back the launches counted while capturing)."""
    global _launches
    _launches += int(n)


def call(name, *args):
    global _launches
    prof = _profile
    pname = name
    if prof is not None and pname in prof:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        flops, tag = 0.0, name
        a0 = args[0] if args else None
        if isinstance(a0, GemmArgs):
            flops = 2.0 * a0.m * a0.n * a0.k
            tag = f"gemm a_mn={a0.a_mn_major} b_mn={a0.b_mn_major} epi={a0.epilogue}"
            if _profile_shapes:
                tag += f" M={a0.m} N={a0.n} K={a0.k}"
        elif isinstance(a0, AttentionFwdArgs):      # full Sq x Sk (bench.py / BASELINE.md convention), head_dim 64
            flops = 4.0 * a0.batch * a0.heads * a0.sq * a0.sk * 64
            tag = f"attention_fwd Sq={a0.sq} Sk={a0.sk} causal={a0.causal} drop={int(a0.drop_p > 0)}"
        elif isinstance(a0, AttentionBwdArgs):
            flops = 10.0 * a0.batch * a0.heads * a0.sq * a0.sk * 64
            tag = f"attention_bwd Sq={a0.sq} Sk={a0.sk} causal={a0.causal} drop={int(a0.drop_p > 0)}"
        prof[pname].append((e0, e1, flops, tag))
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        check(rc, name)
    _launches += _KERNELS_PER_CALL.get(name, 1)


def profile_ops(fn, names, repeats=1):
    """Run fn() `repeats` times with CUDA events around every call of the named entry points (on the launch stream).
    Returns {name: {"ms": device ms per repeat, "calls": launches per repeat, "flops": algorithmic FLOPs per repeat}}."""
    global _profile
    _profile = {n: [] for n in names}
    try:
        for _ in range(repeats):
            fn()
        torch.cuda.synchronize()
        out = {}
        for n, evs in _profile.items():
            detail = {}
            for a, b, f, tag in evs:
                d = detail.setdefault(tag, [0.0, 0.0, 0])
                d[0] += a.elapsed_time(b) / repeats
                d[1] += f / repeats
                d[2] += 1
            out[n] = {"ms": sum(a.elapsed_time(b) for a, b, _, _ in evs) / repeats, "calls": len(evs) // repeats,
                      "flops": sum(f for _, _, f, _ in evs) / repeats,
                      "detail": {t: {"ms": round(v[0], 3), "tflops": round(v[1] / (v[0] * 1e-3) / 1e12, 1) if v[0] > 0 and v[1] > 0 else None,
                                     "calls": v[2] // repeats} for t, v in detail.items()}}
        return out
    finally:
        _profile = None
