"""2+ GPU probe (torchrun): (a) does torch symmetric memory work here, (b) bandwidth of a copy-engine push into a peer's
buffer, (c) how a push / an NCCL all-reduce behaves while persistent tcgen05 GEMMs own every SM of the main stream,
and what it costs those GEMMs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pixparse_b200 import ops
from pixparse_b200.framework import DeviceEnv

env = DeviceEnv()
rank, world, dev = env.global_rank, env.world_size, env.device
torch.cuda.set_device(dev)
def log(*a):
    if rank == 0:
        print(*a, flush=True)

N = 20 * 1024 * 1024      # fp32 elements = 80 MB
peer = (rank + 1) % world
src = torch.full((N,), float(rank + 1), device=dev)

# ---- (a) symmetric memory ------------------------------------------------------------------------------------
sm_ok = False
try:
    import torch.distributed._symmetric_memory as symm
    buf = symm.empty(N, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    peer_buf = hdl.get_buffer(peer, (N,), torch.float32)
    buf.zero_()
    dist.barrier()
    peer_buf.copy_(src)
    torch.cuda.synchronize()
    dist.barrier()
    want = float((rank - 1) % world + 1)
    sm_ok = bool((buf == want).all().item())
    log(f"symmetric memory: alloc + rendezvous + peer write OK={sm_ok} (multicast={getattr(hdl, 'multicast_ptr', 0) != 0})")
except Exception as e:
    log("symmetric memory FAILED:", repr(e)[:300])

# ---- (a') classic CUDA IPC handle exchange -----------------------------------------------------------------------
ipc_ok = False
try:
    mine = torch.zeros(N, device=dev)
    h = mine.untyped_storage()._share_cuda_()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    ph = hs[peer]
    st = torch.UntypedStorage._new_shared_cuda(*ph)
    peer_t = torch.tensor([], dtype=torch.float32, device=st.device).set_(st, 0, (N,))
    dist.barrier()
    peer_t.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    ipc_ok = bool((mine == float((rank - 1) % world + 1)).all().item())
    log(f"CUDA IPC: peer tensor on {peer_t.device}, write OK={ipc_ok}")
except Exception as e:
    log("CUDA IPC FAILED:", repr(e)[:300])

def timed(fn, reps=5, stream=None):
    s = stream or torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize(); dist.barrier()
    with torch.cuda.stream(s):
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

targets = {}
if sm_ok:
    targets["symm"] = peer_buf
if ipc_ok:
    targets["ipc"] = peer_t
for k, t in targets.items():
    ms = timed(lambda: t.copy_(src, non_blocking=True))
    log(f"push 80 MB to peer via {k}: {ms:.3f} ms = {N * 4 / ms / 1e6:.0f} GB/s (idle GPU)")
ar = torch.zeros(N, device=dev)
ms = timed(lambda: dist.all_reduce(ar))
log(f"NCCL all-reduce 80 MB fp32, {world} ranks: {ms:.3f} ms (idle GPU)")

# ---- (c) under load: 40 encoder-sized GEMMs on the main stream ---------------------------------------------------
M, Nn, K = 32288, 3072, 768
a = torch.randn((M, K), device=dev).bfloat16(); w = torch.randn((Nn, K), device=dev).bfloat16()
out = torch.empty((M, Nn), device=dev, dtype=torch.bfloat16)
def gemms(n=40):
    for _ in range(n):
        ops.gemm(a, w, out=out)
side = torch.cuda.Stream(device=dev)
def under_load(label, comm):
    gemms(4); torch.cuda.synchronize(); dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    gemms(5)
    ev = torch.cuda.Event(); ev.record()
    side.wait_event(ev)
    with torch.cuda.stream(side):
        e[2].record()
        if comm is not None:
            comm()
        e[3].record()
    gemms(35)
    e[1].record()
    torch.cuda.synchronize()
    log(f"{label}: 40 GEMMs {e[0].elapsed_time(e[1]):.3f} ms; comm op {e[2].elapsed_time(e[3]):.3f} ms")
under_load("no comm            ", None)
for k, t in targets.items():
    under_load(f"push 80 MB via {k:4s}", lambda t=t: t.copy_(src, non_blocking=True))
    under_load(f"push 8x80 MB via {k:4s}", lambda t=t: [t.copy_(src, non_blocking=True) for _ in range(8)])
under_load("NCCL all-reduce 80MB", lambda: dist.all_reduce(ar))
under_load("NCCL all-reduce x8  ", lambda: [dist.all_reduce(ar) for _ in range(8)])
under_load("no comm (again)     ", None)
dist.barrier()
