"""Host-side logic that needs no GPU: LR schedule, layer-decay grouping, multi-rank gradient reducer (gloo)."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_cosine_schedule_matches_oracle_restatement():
    from oracle import timm_helpers
    from pixparse_b200.schedule import create_scheduler
    p1 = [torch.nn.Parameter(torch.zeros(3))]
    p2 = [torch.nn.Parameter(torch.zeros(3))]
    o1 = torch.optim.AdamW([{"params": p1, "lr_scale": 0.5}], lr=3e-4)
    o2 = torch.optim.AdamW([{"params": p2, "lr_scale": 0.5}], lr=3e-4)
    s1, _ = create_scheduler(o1, 'cosine', warmup_lr=0.0, warmup_intervals=2, num_intervals=10, updates_per_interval=7)
    s2, _ = timm_helpers.create_scheduler_v2(o2, 'cosine', warmup_lr=0.0, warmup_epochs=2, num_epochs=10,
                                             step_on_epochs=False, updates_per_epoch=7)
    for t in list(range(0, 30)) + [35, 69, 70, 71, 100]:
        s1.step_update(t)
        s2.step_update(t)
        assert o1.param_groups[0]["lr"] == pytest.approx(o2.param_groups[0]["lr"], rel=1e-12, abs=1e-18)
    s1.step_update(14)     # end of warm-up: the cosine is not shifted by the warm-up length
    assert o1.param_groups[0]["lr"] == pytest.approx(0.5 * 0.5 * 3e-4 * (1 + math.cos(math.pi * 14 / 70)))


def test_layer_decay_grouping_matches_oracle_restatement():
    from oracle import timm_helpers
    from pixparse_b200.optim import layer_decay_groups
    m = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.LayerNorm(4), torch.nn.Linear(4, 2))
    ours = layer_decay_groups(m, 0.75)
    ref = timm_helpers.param_groups_layer_decay(m, weight_decay=0.0, layer_decay=0.75)
    assert [(g["lr_scale"], len(g["params"])) for g in ours] == [(g["lr_scale"], len(g["params"])) for g in ref]
    m.pretrained_cfg = {"classifier": "2."}          # a model that does name its head gets 12-tensor chunks
    ours = layer_decay_groups(m, 0.75)
    ref = timm_helpers.param_groups_layer_decay(m, weight_decay=0.0, layer_decay=0.75)
    assert sorted((g["lr_scale"], len(g["params"])) for g in ours) == sorted((g["lr_scale"], len(g["params"])) for g in ref)


def _reducer_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pixparse_b200.framework import DeviceEnv
    from pixparse_b200.reducer import GradReducer
    env = DeviceEnv(device_type="cpu", backend="gloo")
    assert (env.world_size, env.global_rank) == (world, rank)
    n = 1000
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    red = GradReducer(flat, bucket_bytes=4 * 300)
    # backward finishes ranges from the top of the arena downwards; the stem [0, 100) is never announced
    red.begin()
    red.range_ready(700, 1000)
    red.range_ready(400, 700)
    red.range_ready(100, 400)
    red.finish()
    expect = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(flat, expect)
    covered = sorted(red._done)
    ok = ok and covered[0][0] == 0 and covered[-1][1] == n
    # accumulation micro-step: no_sync leaves local gradients untouched
    flat2 = torch.full((n,), float(rank + 1))
    red2 = GradReducer(flat2, bucket_bytes=1 << 20)
    with red2.no_sync():
        red2.begin()
        red2.range_ready(0, n)
        red2.finish()
    ok = ok and torch.equal(flat2, torch.full((n,), float(rank + 1)))
    red2.begin()
    red2.range_ready(0, n)
    red2.finish()
    ok = ok and torch.allclose(flat2, torch.full((n,), (1 + world) / 2.0))
    obj = env.broadcast_object({"x": 42} if rank == 0 else None)
    ok = ok and obj == {"x": 42}
    # the copy-engine exchange (reduce-scatter + all-gather by pushes into peer buffers) over the exchange test double:
    # ragged buckets (shares of 64-element granularity, the last ones short or empty), an unannounced stem, no_sync
    from pixparse_b200.reducer import P2PGradReducer, _ExchangeTransport

    class _Arena:
        def __init__(self, g):
            self.g32 = g

        def replace_grad_buffer(self, new):
            new.copy_(self.g32)
            self.g32 = new
    n3 = 64 * 37 + 5
    g = torch.Generator().manual_seed(11)
    base = torch.randn(n3, generator=g)
    arena = _Arena(base * (rank + 1) + rank)
    red3 = P2PGradReducer(arena, bucket_bytes=4 * 700, transport=_ExchangeTransport)
    ok = ok and arena.g32 is red3.flat
    c, shares = red3.shares(64, 64 + 100)
    ok = ok and c == 64 and shares == [(64, 128), (128, 164)][:world] + [(164, 164)] * (world - 2)
    red3.begin()
    red3.range_ready(64 * 30, n3)
    red3.range_ready(64 * 9, 64 * 30)
    red3.range_ready(64 * 8, 64 * 9)
    red3.range_ready(64, 64 * 8)
    red3.finish()
    expect3 = base * (sum(range(1, world + 1)) / world) + sum(range(world)) / world
    ok = ok and torch.allclose(arena.g32, expect3, atol=1e-6)
    before = arena.g32.clone()
    with red3.no_sync():
        red3.begin()
        red3.range_ready(0, n3)
        red3.finish()
    ok = ok and torch.equal(arena.g32, before)
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_grad_reducer_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_reducer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]


def test_synthetic_batch_layout():
    from pixparse_b200 import synthetic
    img, text, tgt = synthetic.synthetic_batch(3, (64, 48), 33, seed=1)
    assert img.shape == (3, 1, 64, 48) and img.dtype == torch.float32 and img.min() >= -1 and img.max() <= 1
    assert (text[:, 0] == synthetic.S_PRETRAIN_ID).all() and (tgt[:, 0] == -100).all()
    for b in range(3):
        eos = (text[b] == synthetic.EOS_ID).nonzero()[0, 0].item()
        assert (text[b, eos + 1:] == synthetic.PAD_ID).all() and (tgt[b, eos + 1:] == -100).all()
        assert (tgt[b, 1:eos + 1] == text[b, 1:eos + 1]).all()
    img2, _, _ = synthetic.synthetic_batch(3, (64, 48), 33, seed=1)
    assert torch.equal(img, img2)
    pages = synthetic.synthetic_pages_u8(2, 110, 85, seed=0)
    assert pages.dtype == torch.uint8 and pages.shape == (2, 110, 85)


def test_cer_wer_edit_distance():
    from pixparse_b200.ocr_utils import cer, wer
    assert cer(["hello world"], ["hello world"]) == 0.0
    assert cer(["abcd"], ["abed"]) == pytest.approx(0.25)
    assert cer(["abcd"], [""]) == pytest.approx(1.0)
    assert wer(["the quick brown fox"], ["the quick fox"]) == pytest.approx(0.25)
    assert wer(["a b", "c d"], ["a x", "c d"]) == pytest.approx(0.25)
    assert cer(["<pad>ab  c<pad>"], ["ab c"]) == 0.0       # <pad> removed, whitespace collapsed / stripped


# ---------------------------------------------------------------------------------------------------------------------
# loader / sample boundary: LoaderBundle.set_interval, sample layouts, the app/train.py interval loop
# ---------------------------------------------------------------------------------------------------------------------
def test_loader_bundle_layouts_and_set_interval():
    from pixparse_b200 import data
    b = data.create_synthetic_loader("pretrain", batch_size=2, num_samples=5, image_size=(32, 32), text_len=9, seed=3)
    assert (b.num_batches, b.num_samples, b.sampler) == (2, 5, None)        # drop_last on the train split
    batches = list(b.loader)
    assert len(batches) == 2
    img, txt, tgt = batches[0]
    assert img.shape == (2, 1, 32, 32) and txt.shape == (2, 9) and tgt.shape == (2, 9) and txt.dtype == torch.int64
    assert (tgt[:, 0] == -100).all()
    b.set_interval(4)
    assert b.loader.dataset.interval == 4
    # distributed: a DistributedSampler is created and set_interval re-seeds its shuffle (data/loader.py:95-104)
    b0 = data.create_synthetic_loader("pretrain", 2, 8, image_size=(32, 32), text_len=9, world_size=2, global_rank=0)
    b1 = data.create_synthetic_loader("pretrain", 2, 8, image_size=(32, 32), text_len=9, world_size=2, global_rank=1)
    assert b0.num_samples == 4 and b0.num_batches == 2
    b0.set_interval(1); b1.set_interval(1)
    i0, i1 = list(b0.sampler), list(b1.sampler)
    assert sorted(i0 + i1) == list(range(8))                                # ranks partition the interval's samples
    b0.set_interval(2)
    assert list(b0.sampler) != i0                                           # another interval, another order
    # eval-OCR layout: lists of lists (task_cruller_eval_ocr.py:199-207)
    e = data.create_synthetic_loader("eval_ocr", 3, 3, image_size=(32, 32), text_len=7, is_train=False)
    img, texts, targets = next(iter(e.loader))
    assert img.shape == (3, 1, 32, 32) and len(targets) == 3 and isinstance(targets[0], list)
    assert torch.stack([t[0] for t in targets]).shape == (3, 7)
    # on a CPU device the prefetcher is transparent
    p = data.create_synthetic_loader("pretrain", 2, 4, image_size=(32, 32), text_len=9, device="cpu")
    assert isinstance(p.loader, data.DevicePrefetcher) and len(list(p.loader)) == 2


def test_train_loop_drives_task_like_app_train(tmp_path):
    """app/train.py:48-67: for each interval set_interval(i), train_one_interval (interval_start, train_step per
    batch, interval_end), then a checkpoint of task.model.state_dict() on the primary rank."""
    from pixparse_b200 import data

    class Env:
        def is_primary(self):
            return True

    class FakeTask:
        def __init__(self):
            self.start_interval, self.num_intervals, self.device_env = 1, 3, Env()
            self.model = torch.nn.Linear(2, 2)
            self.log = []

        def train_interval_start(self):
            self.log.append("start")

        def train_step(self, sample):
            self.log.append(("step", tuple(sample[0].shape)))
            return {}

        def train_interval_end(self):
            self.log.append("end")

    calls = []
    bundle = data.create_synthetic_loader("pretrain", 2, 4, image_size=(32, 32), text_len=9)
    orig = bundle.set_interval
    bundle.set_interval = lambda i: (calls.append(i), orig(i))
    task = FakeTask()
    data.train(task, {"train": bundle}, output_checkpoint_dir=str(tmp_path), experiment="exp")
    assert calls == [1, 2]
    per_interval = ["start", ("step", (2, 1, 32, 32)), ("step", (2, 1, 32, 32)), "end"]
    assert task.log == per_interval * 2
    assert sorted(os.listdir(tmp_path / "exp")) == ["checkpoint-1.pt", "checkpoint-2.pt"]
    sd = torch.load(tmp_path / "exp" / "checkpoint-2.pt")
    assert set(sd) == {"weight", "bias"}
