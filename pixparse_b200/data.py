"""Loader / sample boundary of the train step.

The reference builds its loaders with chug / webdataset (``pixparse.data.create_loader``,
/root/reference/src/pixparse/data/loader.py:24-118) and hands ``train()`` a ``chug.common.LoaderBundle`` per split:
``app/train.py:48-67`` calls ``loader.set_interval(i)`` and then ``train_one_interval(task, loader)``, which iterates
``loader.loader`` (framework/train.py:5-14). chug and webdataset are out of scope here; what the hot path needs from
that layer is restated:

* :class:`LoaderBundle` -- same fields (``loader, num_batches, num_samples, sampler``) and ``set_interval``;
* :class:`SyntheticPages` -- map-style dataset of seeded synthetic samples in the exact layouts the three tasks consume
  (pretrain 3-tuple, RVL-CDIP dict, eval-OCR lists-of-lists), so ``DataLoader`` + ``DistributedSampler`` + the task's
  ``collate_fn`` can be driven as the reference drives them;
* :class:`DevicePrefetcher` -- double-buffered host -> device staging: batch i+1 (pinned) is copied on a side stream
  while batch i trains, so ``train_step`` never waits for PCIe. At ~800 pages/s/GPU the reference's synchronous
  ``.to(device, non_blocking=True)`` from pageable memory (task_cruller_pretrain.py:240-242) would be exposed;
* :func:`train` -- the interval loop of ``app/train.py:48-67`` (set_interval, train_one_interval, checkpoint of
  ``task.model.state_dict()`` on the primary rank).
"""
import os
from collections import deque
from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler

from . import synthetic
from .framework import train_one_interval


@dataclass
class LoaderBundle:
    """chug.common.LoaderBundle as pixparse uses it (data/loader.py:113-118, app/train.py:57)."""
    loader: Any
    num_batches: int = 0
    num_samples: int = 0
    sampler: Optional[Any] = None

    def set_interval(self, interval: int):
        """New interval (epoch): re-seed the shuffling. DistributedSampler -> set_epoch; loaders / datasets that carry
        their own notion of an interval (webdataset shared epoch in chug) -> set_interval / set_epoch."""
        if self.sampler is not None and hasattr(self.sampler, "set_epoch"):
            self.sampler.set_epoch(interval)
        inner = getattr(self.loader, "loader", self.loader)      # see through a DevicePrefetcher
        for obj in (inner, getattr(inner, "dataset", None)):
            if obj is None or obj is self.sampler:
                continue
            if hasattr(obj, "set_interval"):
                obj.set_interval(interval)
            elif hasattr(obj, "set_epoch"):
                obj.set_epoch(interval)

    def __len__(self):
        return self.num_batches


class SyntheticPages(Dataset):
    """Seeded synthetic samples, one page each, in the layout a task's loader yields BEFORE collation.

    kind='pretrain': (image (1,H,W) f32 normalised, text (Lt,) i64, target (Lt,) i64)  -> default_collate -> 3-tuple
    kind='rvlcdip' : {'image': PIL 'L' page, 'label': int}                             -> task.collate_fn   -> dict
    kind='eval_ocr': (image, [text], [target])                                         -> lists of lists
                     (task_cruller_eval_ocr.py:199-207 stacks item[0] of every entry)
    """

    def __init__(self, kind, num_samples, image_size=(576, 448), text_len=513, seed=0, page_size=(1100, 850)):
        assert kind in ("pretrain", "rvlcdip", "eval_ocr")
        self.kind, self.num_samples, self.image_size = kind, int(num_samples), tuple(image_size)
        self.text_len, self.seed, self.page_size = text_len, seed, tuple(page_size)
        self.interval = 0

    def set_interval(self, interval):
        self.interval = int(interval)

    def __len__(self):
        return self.num_samples

    def __getitem__(self, i):
        seed = (self.seed * 1000003 + i) & 0x7FFFFFFF
        if self.kind == "rvlcdip":
            from PIL import Image
            page = synthetic.synthetic_pages_u8(1, self.page_size[0], self.page_size[1], seed=seed)[0]
            return {"image": Image.fromarray(page.numpy(), mode="L"), "label": seed % 16}
        image, text, target = synthetic.synthetic_batch(1, self.image_size, self.text_len, seed=seed)
        if self.kind == "pretrain":
            return image[0], text[0], target[0]
        return image[0], [text[0]], [target[0]]


def eval_ocr_collate(batch):
    """Keeps the annotation lists as lists (one [tensor] per page), stacks the images."""
    return (torch.stack([b[0] for b in batch]), [b[1] for b in batch], [b[2] for b in batch])


def create_synthetic_loader(kind, batch_size, num_samples, *, image_size=(576, 448), text_len=513, seed=0,
                            world_size=1, global_rank=0, num_workers=0, collate_fn=None, is_train=True,
                            device=None, prefetch=2):
    """The hf_dataset branch of data/loader.py:82-118 over SyntheticPages: DataLoader (+ DistributedSampler when
    world_size > 1, drop_last) wrapped in a LoaderBundle; with ``device`` the loader is double-buffered onto it."""
    ds = SyntheticPages(kind, num_samples, image_size=image_size, text_len=text_len, seed=seed)
    sampler = None
    if world_size > 1:
        sampler = DistributedSampler(ds, rank=global_rank, shuffle=is_train, seed=seed, num_replicas=world_size,
                                     drop_last=True)
    if collate_fn is None and kind == "eval_ocr":
        collate_fn = eval_ocr_collate
    base = DataLoader(ds, batch_size=batch_size, sampler=sampler, num_workers=num_workers, collate_fn=collate_fn,
                      drop_last=is_train, pin_memory=False)
    loader = DevicePrefetcher(base, device, depth=prefetch) if device is not None else base
    n = len(sampler) if sampler is not None else len(ds)
    return LoaderBundle(loader=loader, num_batches=len(base), num_samples=n, sampler=sampler)


def _map_tensors(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map_tensors(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_tensors(v, fn) for v in obj)
    return obj


class DevicePrefetcher:
    """Iterates ``loader`` and yields the same samples with every tensor already on ``device``.

    Up to ``depth`` batches are in flight: each is pinned (if it is not already) and copied with non_blocking=True on a
    dedicated copy stream as soon as the host has it; the consumer's stream waits for that batch's event only. Nested
    tuples / lists / dicts are preserved (pretrain 3-tuple, RVL-CDIP dict, eval-OCR lists of lists). On a CPU ``device``
    (host-logic tests) samples pass through untouched."""

    def __init__(self, loader, device, depth=2):
        self.loader = loader
        self.device = torch.device(device)
        self.depth = max(1, int(depth))
        self._stream = None

    def __len__(self):
        return len(self.loader)

    def _stage(self, sample):
        dev = self.device
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        host = _map_tensors(sample, lambda t: t if (t.is_cuda or t.is_pinned()) else t.pin_memory())
        with torch.cuda.stream(self._stream):
            staged = _map_tensors(host, lambda t: t.to(dev, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return staged, ev, host        # `host` stays referenced until the copy has been consumed

    def __iter__(self):
        if self.device.type != "cuda":
            yield from self.loader
            return
        it = iter(self.loader)
        queue = deque()
        exhausted = False
        while True:
            while not exhausted and len(queue) < self.depth:
                try:
                    queue.append(self._stage(next(it)))
                except StopIteration:
                    exhausted = True
            if not queue:
                return
            staged, ev, _host = queue.popleft()
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            _map_tensors(staged, lambda t: (t.record_stream(cur), t)[1])
            yield staged


def train(task, loaders: Dict[str, LoaderBundle], output_checkpoint_dir=None, experiment="", save=True):
    """The interval loop of pixparse.app.train.train (app/train.py:48-67)."""
    train_loader = loaders["train"]
    for i in range(task.start_interval, task.num_intervals):
        train_loader.set_interval(i)
        train_one_interval(task, train_loader)
        if save and output_checkpoint_dir is not None and task.device_env.is_primary():
            checkpoint_dir = os.path.join(output_checkpoint_dir, experiment)
            os.makedirs(checkpoint_dir, exist_ok=True)
            torch.save(task.model.state_dict(), os.path.join(checkpoint_dir, f"checkpoint-{i}.pt"))
