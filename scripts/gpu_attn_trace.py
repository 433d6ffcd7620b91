"""Timeline of one attention-backward CTA (debug build with -DAB_TRACE): clock64 stamps per query tile."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixparse_b200 import ops, _lib
B, H, S = 32, 12, 1009
D = H * 64
torch.manual_seed(0)
qkv = (torch.randn((B * S, 3 * D), device="cuda") * 0.5).bfloat16()
dout = torch.randn((B * S, D), device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
out, lse = ops.attention_fwd(qkv, qkv, qkv, B=B, H=H, Sq=S, Sk=S, q_col0=0, k_col0=D, v_col0=2 * D)
trace = torch.zeros(8 * 16, device="cuda", dtype=torch.int64)
lib = _lib.lib()
lib.b200_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.b200_debug_set_trace(trace.data_ptr())
for _ in range(3):
    ops.attention_bwd(qkv, qkv, qkv, out, dout, lse, dqkv, dqkv, dqkv, B=B, H=H, Sq=S, Sk=S, q_col0=0, k_col0=D,
                      v_col0=2 * D, dq_col0=0, dk_col0=D, dv_col0=2 * D)
torch.cuda.synchronize()
t = trace.cpu().view(8, 16)
base = int(t[0, 0])
names = ["A:loop_top", "A:s_full", "A:exp0_done", "A:p_ready_arrive", "B:loop_top", "B:p_ready", "B:ds_ready_arrive",
         "B:drain_done", "m:before_p_ready", "m:p_ready", "m:before_ds_ready", "m:ds_ready", "m:issued_all"]
print("cycles relative to compute loop top of tile 0 (CTA 3,0,0; A warp 0 / B warp 4 / MMA warp)")
for it in range(8):
    print(f"tile {it}: " + "  ".join(f"{n}={int(t[it, i]) - base}" for i, n in enumerate(names)))
