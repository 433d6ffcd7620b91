"""Library baseline for the attention kernels: torch F.scaled_dot_product_attention (flash / cuDNN / mem-efficient
backends), forward and forward+backward, at the attention shapes of the Cruller step, next to this repo's kernels at the
same shapes. This is the bar SURVEY 2.3 K5/K9/K10 sets (the reference reaches SDPA through timm Attention and
BartAttention). Prints one JSON object; run on a B200 (scripts only time, they assert nothing).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

from pixparse_b200 import ops

iters = int(os.environ.get("ITERS", "10"))
torch.manual_seed(0)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# name, B, H, Sq, Sk, causal, dropout
SHAPES = [
    ("base encoder  S=1009", 32, 12, 1009, 1009, False, 0.0),
    ("base dec-self T=512 causal p=0.1", 32, 12, 512, 512, True, 0.1),
    ("base dec-cross 512x1009 p=0.1", 32, 12, 512, 1009, False, 0.1),
    ("base dec-self T=512 causal p=0", 32, 12, 512, 512, True, 0.0),
    ("base dec-cross 512x1009 p=0", 32, 12, 512, 1009, False, 0.0),
    ("large encoder S=2509", 8, 16, 2509, 2509, False, 0.0),
]
BACKENDS = [("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION),
            ("efficient", SDPBackend.EFFICIENT_ATTENTION)]

out = {"iters": iters, "flops": "4*B*H*Sq*Sk*64 fwd, 10*... bwd, causal counted FULL (bench.py convention)", "shapes": []}
for name, B, H, Sq, Sk, causal, p in SHAPES:
    D = H * 64
    rec = {"shape": name, "B": B, "H": H, "Sq": Sq, "Sk": Sk, "causal": causal, "dropout": p}
    f_fwd = 4.0 * B * H * Sq * Sk * 64
    f_bwd = 10.0 * B * H * Sq * Sk * 64
    # ---- library: [B, H, S, 64] contiguous (its best case)
    q = (torch.randn((B, H, Sq, 64), device="cuda") * 0.5).bfloat16().requires_grad_(True)
    k = (torch.randn((B, H, Sk, 64), device="cuda") * 0.5).bfloat16().requires_grad_(True)
    v = (torch.randn((B, H, Sk, 64), device="cuda") * 0.5).bfloat16().requires_grad_(True)
    do = torch.randn((B, H, Sq, 64), device="cuda").bfloat16()
    for bname, be in BACKENDS:
        try:
            with sdpa_kernel(be):
                fwd = lambda: F.scaled_dot_product_attention(q, k, v, dropout_p=p, is_causal=causal)
                with torch.no_grad():
                    ms_f = timeit(fwd)

                def fb():
                    o = F.scaled_dot_product_attention(q, k, v, dropout_p=p, is_causal=causal)
                    o.backward(do)
                    q.grad = k.grad = v.grad = None
                ms_fb = timeit(fb)
            rec[bname] = {"fwd_ms": round(ms_f, 4), "fwd_bwd_ms": round(ms_fb, 4), "bwd_ms_est": round(ms_fb - ms_f, 4),
                          "fwd_tflops": round(f_fwd / ms_f / 1e9, 1),
                          "bwd_tflops_est": round(f_bwd / max(ms_fb - ms_f, 1e-6) / 1e9, 1)}
        except Exception as e:      # backend not available for this shape / build
            rec[bname] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
    del q, k, v, do
    # ---- this repo: packed [B*S, width] activations, heads addressed by column offset (no transposes)
    q2 = (torch.randn((B * Sq, D), device="cuda") * 0.5).bfloat16()
    kv2 = (torch.randn((B * Sk, 2 * D), device="cuda") * 0.5).bfloat16()
    do2 = torch.randn((B * Sq, D), device="cuda").bfloat16()
    dq2, dkv2 = torch.empty_like(q2), torch.empty_like(kv2)
    kw = dict(B=B, H=H, Sq=Sq, Sk=Sk, q_col0=0, k_col0=0, v_col0=D, causal=causal, drop=(p, 1234) if p > 0 else None)
    fwd2 = lambda: ops.attention_fwd(q2, kv2, kv2, **kw)
    o2, lse2 = fwd2()
    bwd2 = lambda: ops.attention_bwd(q2, kv2, kv2, o2, do2, lse2, dq2, dkv2, dkv2, dq_col0=0, dk_col0=0, dv_col0=D, **kw)
    ms_f, ms_b = timeit(fwd2), timeit(bwd2)
    rec["b200"] = {"fwd_ms": round(ms_f, 4), "bwd_ms": round(ms_b, 4), "fwd_bwd_ms": round(ms_f + ms_b, 4),
                   "fwd_tflops": round(f_fwd / ms_f / 1e9, 1), "bwd_tflops": round(f_bwd / ms_b / 1e9, 1)}
    out["shapes"].append(rec)
    print(json.dumps(rec), flush=True)
    del q2, kv2, do2, dq2, dkv2, o2, lse2
    torch.cuda.empty_cache()

path = os.environ.get("OUT")
if path:
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
