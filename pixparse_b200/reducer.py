"""Bucketed gradient all-reduce over the flat gradient arena, overlapped with backward.

Stands in for ``torch.nn.parallel.DistributedDataParallel(model, device_ids=[device], static_graph=True)``
(/root/reference/src/pixparse/task/task_cruller_pretrain.py:181-189) and ``model.no_sync()`` (:280-283): gradients
are AVERAGED over ranks, buckets are issued in the order backward finishes them (decoder first, then encoder blocks
from last to first) on a side stream so NCCL runs under the remaining backward kernels, and the optimizer waits
for the last bucket. Because the gradients already live in one contiguous arena there is no bucket copy-in/out.
"""
import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, flat_grads, process_group=None, bucket_bytes=64 << 20):
        # tuning knob: PIXPARSE_B200_REDUCE_BUCKET_MB=<n> (0 = one all-reduce of the whole arena after backward)
        import os
        env = os.environ.get("PIXPARSE_B200_REDUCE_BUCKET_MB")
        if env is not None:
            bucket_bytes = int(env) << 20 if int(env) > 0 else (1 << 62)
        self.flat = flat_grads
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_elems = max(1, bucket_bytes // flat_grads.element_size())
        self.cuda = flat_grads.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat_grads.device) if self.cuda else None
        self.backend = dist.get_backend(process_group) if dist.is_initialized() else None
        self._pending = None       # [lo, hi) finished by backward but not yet issued
        self._done = []            # issued ranges (for bookkeeping / tests)
        self.enabled = True        # False inside no_sync() (gradient accumulation micro-steps)

    # -- context manager mirroring DDP.no_sync()
    class _NoSync:
        def __init__(self, r):
            self.r = r

        def __enter__(self):
            self.prev = self.r.enabled
            self.r.enabled = False

        def __exit__(self, *a):
            self.r.enabled = self.prev

    def no_sync(self):
        return GradReducer._NoSync(self)

    def begin(self):
        self._pending = None
        self._done = []

    def _issue(self, lo, hi):
        if hi <= lo:
            return
        view = self.flat[lo:hi]
        self._done.append((lo, hi))
        if self.world_size == 1:
            return
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                self._all_reduce_mean(view)
        else:
            self._all_reduce_mean(view)

    def _all_reduce_mean(self, view):
        if self.backend == "nccl":
            dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            view.mul_(1.0 / self.world_size)

    def range_ready(self, lo, hi):
        """Backward finished every gradient in arena elements [lo, hi)."""
        if not self.enabled:
            return
        if self._pending is None:
            self._pending = [lo, hi]
        elif hi == self._pending[0]:
            self._pending[0] = lo            # ranges arrive in descending address order
        elif lo == self._pending[1]:
            self._pending[1] = hi
        else:
            self._issue(*self._pending)
            self._pending = [lo, hi]
        if self._pending[1] - self._pending[0] >= self.bucket_elems:
            self._issue(*self._pending)
            self._pending = None

    def finish(self):
        """Issue whatever is left (everything not yet covered) and make the compute stream wait for NCCL."""
        if not self.enabled:
            return
        if self._pending is not None:
            self._issue(*self._pending)
            self._pending = None
        covered = sorted(self._done)
        cur = 0
        gaps = []
        for lo, hi in covered:
            if lo > cur:
                gaps.append((cur, lo))
            cur = max(cur, hi)
        if cur < self.flat.numel():
            gaps.append((cur, self.flat.numel()))
        for lo, hi in gaps:
            self._issue(lo, hi)
        if self.cuda and self.world_size > 1:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)
