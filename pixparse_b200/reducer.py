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
        self.measure_exposed = False
        self._exposed = []
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
        # a full bucket goes out; so does anything that reaches into the first bucket's worth of the arena -- backward ends
        # there (ranges arrive in descending address order), and whatever is still pending at that point is exchanged
        # with nothing left to hide it
        if self._pending[1] - self._pending[0] >= self.bucket_elems or self._pending[0] < self.bucket_elems:
            self._issue(*self._pending)
            self._pending = None

    def exposed_ms(self):
        """Mean device time the compute stream spent waiting for the exchange at the end of backward (events recorded by
        finish() when `measure_exposed` is set; call after a synchronize)."""
        if not self._exposed:
            return None
        ms = [a.elapsed_time(b) for a, b in self._exposed]
        self._exposed = []
        return sum(ms) / len(ms)

    def finish(self):
        """Issue whatever is left (everything not yet covered) and make the compute stream wait for NCCL."""
        if not self.enabled:
            return
        if self.measure_exposed and self.cuda:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self.flat.device))
            self._finish()
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.flat.device))
            self._exposed.append((e0, e1))
            return
        self._finish()

    def _finish(self):
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


# ----------------------------------------------------------------------------------------------------------------------
# Copy-engine gradient exchange over NVLink / NVSwitch
# ----------------------------------------------------------------------------------------------------------------------
class _SymmetricTransport:
    """Peer-addressable buffers through torch symmetric memory (CUDA VMM handles exchanged at rendezvous): a ``copy_``
    into a peer's buffer is a device-to-device memcpy that the COPY ENGINES carry over NVLink -- no SM is involved.

    The barrier is SM-free as well: stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32, executed by the
    GPU's front end) on a row of flags in symmetric memory -- every rank writes the barrier's sequence number into its
    slot of every peer's row, then waits for all slots of its own row to reach it. A write-value is ordered after the
    copies that precede it on the stream, so a rank that sees a peer's number also sees that peer's pushes. (A collective
    kernel for the barrier cannot share an SM with the step's persistent GEMM CTAs, which own the whole register file:
    each of the ~25 barriers of a step then holds back one GEMM CTA, i.e. the whole statically scheduled GEMM, for its
    duration. PIXPARSE_B200_P2P_BARRIER=nccl keeps the 1-element NCCL all-reduce.)"""

    def __init__(self, numel_by_name, device, group):
        import os
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group = group if group is not None else dist.group.WORLD
        self.local, self.peer = {}, {}
        for name, numel in numel_by_name.items():
            t = symm.empty(numel, dtype=torch.float32, device=device)
            t.zero_()
            hdl = symm.rendezvous(t, self.group)
            self.local[name] = t
            self.peer[name] = [t if r == self.rank else hdl.get_buffer(r, (numel,), torch.float32) for r in range(self.world)]
        self._flag = torch.zeros(1, device=device)
        self._drv = None
        if os.environ.get("PIXPARSE_B200_P2P_BARRIER", "memop") == "memop":
            try:
                from cuda.bindings import driver as drv
            except ImportError:
                try:
                    from cuda import cuda as drv
                except ImportError:
                    drv = None
            if drv is not None:
                flags = symm.empty(64, dtype=torch.int32, device=device)      # slot s: written by rank s
                flags.zero_()
                fh = symm.rendezvous(flags, self.group)
                self._flags = flags
                self._flag_ptrs = [int(p) for p in fh.buffer_ptrs]
                self._seq = 0
                self._drv = drv
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)      # every rank's flags are zeroed before anyone writes one
        if self._drv is not None:
            # one round trip now: a driver / device without stream memory operations on peer memory must show up here, on
            # every rank alike, not in the middle of a step
            ok = torch.ones(1, device=device)
            try:
                self.barrier()
                torch.cuda.synchronize(device)
            except RuntimeError:
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if ok.item() == 0:
                self._drv = None

    def push(self, peer, name, offset, src):
        self.peer[name][peer][offset:offset + src.numel()].copy_(src, non_blocking=True)

    def barrier(self):
        """Stream-ordered: every rank's earlier pushes (same stream, in order) have landed once the stream gets past this."""
        if self._drv is None:
            dist.all_reduce(self._flag, group=self.group)      # tiny NCCL all-reduce on the current stream
            return
        drv = self._drv
        self._seq = (self._seq + 1) & 0x7FFFFFFF
        stream = torch.cuda.current_stream().cuda_stream
        me = self.rank
        for d in range(1, self.world):
            p = (me + d) % self.world
            (err,) = drv.cuStreamWriteValue32(stream, self._flag_ptrs[p] + 4 * me, self._seq, 0)
            if int(err) != 0:
                raise RuntimeError(f"cuStreamWriteValue32 failed: {err}")
        for d in range(1, self.world):
            p = (me + d) % self.world
            (err,) = drv.cuStreamWaitValue32(stream, self._flag_ptrs[me] + 4 * p, self._seq, 0)      # cyclic >=
            if int(err) != 0:
                raise RuntimeError(f"cuStreamWaitValue32 failed: {err}")


class _ExchangeTransport:
    """Test double for world_size > 1 WITHOUT peer memory (gloo on CPU): pushes are queued and delivered at the next
    barrier through an object all-gather. Exercises the bucket / share arithmetic and the ordering of the protocol."""

    def __init__(self, numel_by_name, device, group):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group = group
        self.local = {name: torch.zeros(numel, device=device) for name, numel in numel_by_name.items()}
        self._queue = []

    def push(self, peer, name, offset, src):
        self._queue.append((peer, name, offset, src.detach().cpu().clone()))

    def barrier(self):
        everyone = [None] * self.world
        dist.all_gather_object(everyone, self._queue, group=self.group)
        self._queue = []
        for q in everyone:
            for peer, name, offset, data in q:
                if peer == self.rank:
                    self.local[name][offset:offset + data.numel()].copy_(data)


class P2PGradReducer(GradReducer):
    """Gradient mean over ranks WITHOUT collective kernels: what DistributedDataParallel's bucketed all-reduce
    (/root/reference/src/pixparse/task/task_cruller_pretrain.py:181-189) computes, moved onto the copy engines.

    Why: every GEMM / attention kernel of the step is persistent and owns all 148 SMs. An NCCL all-reduce needs SMs of its
    own, so it either waits for kernel boundaries or takes SMs away from kernels whose tiles were dealt out for 148 of them
    (measured on 2 B200s: 640 MB of all-reduces under a stream of GEMMs cost the GEMMs +16 %; the same bytes pushed by the
    copy engines +1.5 % -- profiles/r02_p2p_probe.txt). The exchange therefore uses the one resource backward leaves idle:
      1. each bucket [lo, hi) of the flat gradient arena is cut into `world` shares; rank r owns share r,
      2. every rank PUSHES share p of its gradients into a staging slot on peer p (copy engine, NVLink),
      3. after a barrier the owner sums its share with the `world - 1` staged copies and scales by 1 / world
         (b200_reduce_shards: one small memory-bound kernel),
      4. the owner pushes the finished share into every peer's gradient arena (copy engine again),
    i.e. reduce-scatter + all-gather, per bucket, in the order backward finishes the buckets, on a side stream. Only the
    barriers (1-element NCCL all-reduces) and the local sum touch an SM. Every rank ends with bit-identical averaged
    gradients (each share is summed on exactly one rank), so clipping and AdamW stay replicated and unchanged."""

    def __init__(self, arena, process_group=None, bucket_bytes=64 << 20, transport=None):
        super().__init__(arena.g32, process_group, bucket_bytes)
        assert self.world_size > 1
        self.rank = dist.get_rank(process_group)
        total = arena.g32.numel()
        # staging area: per bucket `world` slots of one share each (shares are rounded up to 64 elements)
        self.stage_numel = total + self.world_size * 64 * 512
        sizes = {"grad": total, "stage": self.stage_numel}
        if transport is None:
            transport = _SymmetricTransport if self.cuda else _ExchangeTransport
        self.tr = transport(sizes, arena.g32.device, process_group)
        # the gradient arena itself must be peer-addressable (step 4 writes into it): move it into the symmetric buffer
        arena.replace_grad_buffer(self.tr.local["grad"])
        self.flat = arena.g32
        self.stage = self.tr.local["stage"]
        self._cursor = 0
        import os
        self._dry = os.environ.get("PIXPARSE_B200_P2P_DRY", "0") == "1"
        if self.cuda:
            # the per-bucket sum kernel should take the first SMs a finishing GEMM frees (its partly empty last wave)
            # instead of queueing behind the next persistent kernel
            self.comm_stream = torch.cuda.Stream(device=arena.g32.device, priority=-1)

    def begin(self):
        super().begin()
        self._cursor = 0

    def shares(self, lo, hi):
        """[(start, stop)] * world: the share of bucket [lo, hi) each rank owns (64-element granularity, may be empty)."""
        w = self.world_size
        c = ((hi - lo + w - 1) // w + 63) // 64 * 64
        return c, [(min(lo + r * c, hi), min(lo + (r + 1) * c, hi)) for r in range(w)]

    def _exchange(self, lo, hi):
        from . import ops
        if self._dry:      # diagnostics: the symmetric-memory arena without any exchange
            return
        w, me = self.world_size, self.rank
        c, shares = self.shares(lo, hi)
        base = self._cursor
        self._cursor += w * c
        if self._cursor > self.stage_numel:
            raise RuntimeError("gradient staging area exhausted (more than 512 buckets in one step?)")
        for d in range(1, w):                      # staggered, so that at any moment every peer is some rank's target
            p = (me + d) % w
            a, b = shares[p]
            if b > a:
                self.tr.push(p, "stage", base + me * c, self.flat[a:b])
        self.tr.barrier()
        a, b = shares[me]
        if b > a:
            if self.cuda:
                ops.reduce_shards(self.flat[a:b], self.stage[base:base + w * c], c, w, me, 1.0 / w)
            else:
                acc = self.flat[a:b].clone()
                for s in range(w):
                    if s != me:
                        acc += self.stage[base + s * c: base + s * c + (b - a)]
                self.flat[a:b].copy_(acc / w)
            for d in range(1, w):
                self.tr.push((me + d) % w, "grad", a, self.flat[a:b])

    def _all_reduce_mean(self, view):
        lo = view.storage_offset() - self.flat.storage_offset()
        self._exchange(lo, lo + view.numel())

    def _finish(self):
        super()._finish()
        if self.world_size > 1 and not self._dry:
            # all step-4 pushes into this rank's arena have landed before the optimizer reads it
            if self.cuda:
                with torch.cuda.stream(self.comm_stream):
                    self.tr.barrier()
                torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)
            else:
                self.tr.barrier()


def make_grad_reducer(arena, process_group=None):
    """The gradient exchange of a data-parallel run: copy-engine pushes over NVLink on B200s (PIXPARSE_B200_REDUCER=p2p,
    the default with NCCL on CUDA), NCCL all-reduce of arena ranges otherwise (PIXPARSE_B200_REDUCER=nccl, gloo, CPU)."""
    import os
    mode = os.environ.get("PIXPARSE_B200_REDUCER", "p2p")
    if mode == "none":      # diagnostics only: independent replicas, no exchange (the floor of the N-GPU step time)
        return None
    if mode == "p2p" and arena.g32.is_cuda and dist.get_backend(process_group) == "nccl":
        try:
            return P2PGradReducer(arena, process_group)
        except (RuntimeError, ImportError, AttributeError) as e:      # no peer-addressable memory on this system
            import warnings
            warnings.warn(f"pixparse_b200: symmetric memory unavailable ({e!r:.200}); gradient exchange falls back to "
                          "NCCL all-reduce")
    return GradReducer(arena.g32, process_group)
