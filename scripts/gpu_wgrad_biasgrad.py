"""Weight-gradient GEMMs of the cruller_base step with and without the fused bias gradient (B200GemmArgs.bias_grad), next
to the stand-alone column-sum kernel they replace. usage: [PIXPARSE_B200_LIB=...] python scripts/gpu_wgrad_biasgrad.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import ops, _lib

torch.manual_seed(0)
M = 32 * 1009
D, F = 768, 3072


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("lib:", _lib.LIB_PATH)
tot = [0.0, 0.0, 0.0]
for name, tokens, n_out, n_in, count in [("enc fc1", M, F, D, 12), ("enc fc2", M, D, F, 12), ("enc qkv", M, 3 * D, D, 12),
                                          ("enc proj", M, D, D, 12), ("dec fc1", 16384, F, D, 4), ("dec kv", M, 2 * D, D, 4),
                                          ("dec 768", 16384, D, D, 12)]:
    dy = torch.randn((tokens, n_out), device="cuda").bfloat16()
    x = torch.randn((tokens, n_in), device="cuda").bfloat16()
    dw = torch.zeros((n_out, n_in), device="cuda")
    db = torch.zeros((n_out,), device="cuda")
    t_plain = timeit(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, epi=ops.EPI_REDUCE_F32, out=dw))
    t_fused = timeit(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, epi=ops.EPI_REDUCE_F32, out=dw, bias_grad=db))
    t_col = timeit(lambda: ops.colsum(dy, db))
    fl = 2.0 * tokens * n_out * n_in
    tot[0] += t_plain * count; tot[1] += t_fused * count; tot[2] += t_col * count
    print(f"{name:9s} tokens={tokens} {n_out}x{n_in}: wgrad {t_plain:7.1f} us ({fl / t_plain / 1e6:6.0f} TF/s)  +bias_grad {t_fused:7.1f} us "
          f"({100 * (t_fused / t_plain - 1):+5.1f} %)  colsum alone {t_col:6.1f} us", flush=True)
print(f"per step: plain {tot[0] / 1e3:.2f} ms, fused {tot[1] / 1e3:.2f} ms, plain + colsum {(tot[0] + tot[2]) / 1e3:.2f} ms")
