"""Multi-GPU check (torchrun, NCCL): a W-rank step on W x B pages must reproduce the 1-rank step on the concatenated
batch when every rank sees the same number of valid target tokens (DDP averages per-rank mean losses unweighted:
SURVEY 8e). Also reports the overlap structure of the gradient reducer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pixparse_b200 import models, synthetic
from pixparse_b200.engine import engine_for
from pixparse_b200.framework import DeviceEnv
from pixparse_b200.reducer import GradReducer, P2PGradReducer

env = DeviceEnv()
rank, world, dev = env.global_rank, env.world_size, env.device
torch.cuda.set_device(dev)
name = os.environ.get("MODEL", "cruller_test")


def build():
    cfg = models.get_model_config(name)
    cfg.image_encoder.pretrained = cfg.text_decoder.pretrained = False
    torch.manual_seed(0)
    m = models.Cruller(cfg)
    m.text_decoder.trunk.resize_token_embeddings(synthetic.PRETRAIN_VOCAB)
    m.text_decoder.trunk.set_dropout(0.0)
    return m.to(dev), cfg


B = int(os.environ.get("B", "4"))
m, cfg = build()
size = tuple(cfg.image_encoder.image_size)
# full batch: no padding so that every rank has the same number of valid tokens
g = torch.Generator().manual_seed(7)
image = (torch.rand((world * B, 1) + size, generator=g) - 0.5) / 0.5
text = torch.randint(3, 50265, (world * B, 33), generator=g)
text[:, 0] = synthetic.S_PRETRAIN_ID
target = text.clone(); target[:, 0] = -100
sl = slice(rank * B, (rank + 1) * B)

eng = engine_for(m)
arena = eng.ensure_bound()
dist.broadcast(arena.p32, src=0)
mode = os.environ.get("PIXPARSE_B200_REDUCER", "p2p")
red = P2PGradReducer(arena, bucket_bytes=1 << 20) if mode == "p2p" else GradReducer(arena.g32, bucket_bytes=1 << 20)
def ready(first, last, _ar=arena):
    lo = _ar.index[first][0]; o, n, _ = _ar.index[last]
    red.range_ready(lo, o + (n + 63) // 64 * 64)
eng._grad_ready_hook = ready
eng.zero_grads()
red.begin()
stats = eng.forward_backward(image[sl].to(dev), text[sl, :-1].contiguous().to(dev), target[sl, 1:].contiguous().to(dev))
red.finish()
torch.cuda.synchronize()
loss_local = stats[1].clone()
dist.all_reduce(loss_local); loss_mean = loss_local.item() / world
g_ddp = arena.g32.clone()

ok = True
if rank == 0:
    m1, _ = build()
    eng1 = engine_for(m1)
    eng1.ensure_bound().p32.copy_(arena.p32)
    eng1.zero_grads()
    s1 = eng1.forward_backward(image.to(dev), text[:, :-1].contiguous().to(dev), target[:, 1:].contiguous().to(dev))
    torch.cuda.synchronize()
    g1 = eng1.arena.g32
    rel = ((g_ddp - g1).norm() / g1.norm()).item()
    dl = abs(loss_mean - s1[1].item()) / s1[1].item()
    ok = rel < 2e-3 and dl < 1e-5
    print(f"ddp_check world={world} model={name} reducer={type(red).__name__}: loss {loss_mean:.6f} vs single {s1[1].item():.6f} (rel {dl:.2e}); "
          f"grad rel-L2 diff {rel:.3e}; reducer issued {len(red._done)} ranges -> {'PASS' if ok else 'FAIL'}", flush=True)
# every rank must hold identical averaged gradients
chk = torch.stack([g_ddp.double().sum(), g_ddp.double().abs().sum(), (g_ddp.double() * torch.arange(g_ddp.numel(), device=dev) % 7).sum()])
lst = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(lst, chk)
same = all(torch.equal(x, lst[0]) for x in lst) if mode == "p2p" else \
    all(((x - lst[0]).abs() <= 1e-9 * lst[0].abs().clamp(min=1.0)).all().item() for x in lst)
if rank == 0:
    print("identical gradients on all ranks:", same, flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
