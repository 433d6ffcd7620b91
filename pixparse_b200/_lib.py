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
        _lib.b200_last_error.restype = ctypes.c_char_p
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

# name -> argtypes; must mirror include/pixparse_b200.h exactly (tests/test_abi.py checks the symbol list)
SIGNATURES = {
    "b200_abi_version": [],
    "b200_device_check": [],
    "b200_debug_gemm_desc": [_I, _I, _I, _I, _I, _I],
    "b200_debug_gemm_single_cta": [_I],
    "b200_debug_attention_bwd_query_major": [_I],
    "b200_attention_fwd": [_P, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "b200_attention_fwd_strided": [_P, _L, _L, _I, _P, _L, _L, _I, _P, _L, _L, _I, _P, _L, _L, _P, _I, _I, _I, _I, _I, _I,
                                   _F, _P],
    "b200_attention_bwd": [_P, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _P, _L, _I, _P, _P, _L, _I, _P, _L, _I, _P, _L,
                           _I, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "b200_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "b200_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "b200_colsum_bf16": [_P, _L, _I, _I, _P, _P],
    "b200_patch_unfold": [_P, _P, _I, _I, _I, _I, _I, _L, _P],
    "b200_tokens_assemble": [_P, _P, _P, _P, _I, _I, _I, _P],
    "b200_tokens_assemble_bwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "b200_embed_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "b200_embed_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _L, _P],
    "b200_cast_f32_bf16": [_P, _P, _L, _P],
    "b200_ce_prepare": [_P, _I, _L, _P, _P],
    "b200_ce_fwd_bwd": [_P, _L, _P, _P, _L, _P, _P, _I, _I, _L, _F, _P],
    "b200_grad_norm": [_P, _L, _P, _P, _F, _F, _P],
    "b200_grad_norm_workspace_floats": [],
    "b200_adamw_step": [_P, _P, _P, _P, _P, _L, _P, _I, _P, _F, _F, _F, _F, _F, _I, _I, _P],
    "b200_preprocess_pages": [_P, _I, _I, _I, _L, _P, _I, _I, _F, _F, _P, _P],
    "b200_gemm_bf16_dropout": [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _P, _L, _P, _L, _P, _P, _L, _I, _I, _F, _U, _P],
    "b200_layernorm_fwd_dropout": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _U, _P],
    "b200_layernorm_bwd_dropout": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _U, _F, _U, _P],
    "b200_attention_fwd_dropout": [_P, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _P, _I, _I, _I, _I, _I, _I, _F, _F, _U, _P],
    "b200_attention_bwd_dropout": [_P, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _P, _L, _I, _P, _P, _L, _I, _P, _L, _I, _P,
                                   _L, _I, _P, _I, _I, _I, _I, _I, _I, _F, _F, _U, _P],
    "b200_gemm_bf16": [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _P, _L, _P, _L, _P, _P, _L, _I, _I, _P],
}


def _declare(l):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    l.b200_attention_bwd_workspace_bytes.argtypes = [_I, _I, _I]
    l.b200_attention_bwd_workspace_bytes.restype = ctypes.c_longlong
    l.b200_preprocess_workspace_bytes.argtypes = [_I, _I]
    l.b200_preprocess_workspace_bytes.restype = ctypes.c_longlong


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
_KERNELS_PER_CALL = {"b200_grad_norm": 2, "b200_attention_bwd": 3, "b200_attention_bwd_dropout": 3}
_launches = 0
_profile = None     # {name: [(start_event, end_event, flops), ...]} while profile_ops() is active


def reset_launch_count():
    global _launches
    _launches = 0


def launch_count():
    return _launches


def call(name, *args):
    global _launches
    prof = _profile
    pname = name[:-8] if name.endswith("_dropout") else name      # dropout variants are profiled with their base op
    if prof is not None and pname in prof:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        is_gemm = name.startswith("b200_gemm_bf16")
        flops = 2.0 * args[6] * args[7] * args[8] if is_gemm else 0.0
        tag = (f"gemm a_mn={args[2]} b_mn={args[5]} epi={args[9]}" if is_gemm else name)
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
