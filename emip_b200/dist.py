"""Multi-GPU plumbing for the hot path: batch sharding and the training gradient all-reduce.

The path shards by frame pairs / clips with no data-path collective.  The one collective of the
reference is DDP's gradient all-reduce (train.py:279); ``FlatGradAllReduce`` is the explicit equivalent for the
learnable prompt / decoder parameters: gradients are packed into ONE flat fp32 buffer and reduced with ONE
``all_reduce`` (NCCL over NVLink on the GPU box, gloo in the CPU tests), which replaces DDP's
``find_unused_parameters=True`` graph walk; ``BucketedGradAllReduce`` is the overlapped form (buckets reduced on a side
stream while the backward pass is still running).
"""
import torch
import torch.distributed as dist


def shard_batch(n_items, world_size, rank):
    """Contiguous [start, stop) slice of ``n_items`` frame pairs / clips owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FlatGradAllReduce:
    """Average the gradients of ``params`` across the process group through one flat buffer (blocking form).

    A parameter that received no gradient on ANY rank keeps ``grad = None`` (what DDP with ``find_unused_parameters=True``
    does, so AdamW creates no state for it and applies no weight decay -- the reference leaves ``adaptor_fc1/2`` trainable
    but unused, train.py:340-342); one that is unused on some ranks only contributes zeros there.
    """

    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.numel = sum(p.numel() for p in self.params)
        if not self.params:
            raise ValueError("no trainable parameters")
        p0 = self.params[0]
        # the per-parameter "used" flags ride behind the gradients in the same buffer: one collective
        self.flat = torch.zeros(self.numel + len(self.params), dtype=torch.float32, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.used = self.flat[self.numel:]

    @property
    def nbytes(self):
        return self.numel * 4

    def __call__(self):
        """Pack -> all_reduce(SUM) -> divide by world size -> unpack into ``p.grad``."""
        for i, (p, v) in enumerate(zip(self.params, self.views)):
            if p.grad is None:
                v.zero_()
                self.used[i] = 0.0
            else:
                v.copy_(p.grad)
                self.used[i] = 1.0
        world = 1
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            world = dist.get_world_size(self.group)
        used = self.used.tolist()
        if world > 1:
            self.flat[:self.numel].div_(world)
        for p, v, u in zip(self.params, self.views, used):
            if u == 0:
                p.grad = None
            elif p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


class BucketedGradAllReduce:
    """Gradient all-reduce overlapped with the backward pass (what DDP does for the reference, train.py:279, with our own
    bookkeeping): the trainable parameters live as views of ONE flat fp32 gradient buffer cut into buckets in reverse
    registration order (the order gradients become ready); a post-accumulate hook per parameter counts its bucket down and,
    when the bucket is complete, launches ``all_reduce`` on it from a side stream while the backward kernels keep running on
    the main stream.  ``finish()`` joins the side stream and scales by 1 / world.  ``p.grad`` IS the bucket view
    (no pack / unpack copies).

    Statistics of the last step: ``comm_ms`` (sum of the all-reduce durations on the side stream), ``exposed_ms`` (how long the
    main stream had to wait in ``finish()``), ``nbytes``.
    """

    def __init__(self, params, bucket_mb=25.0, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self.group = process_group
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        order = list(reversed(self.params))                    # gradients arrive roughly in reverse registration order
        cap = max(1, int(bucket_mb * 1e6 / 4))
        self.buckets, cur, cur_n, off = [], [], 0, 0
        self._where = {}
        for p in order:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            cur.append(p)
            cur_n += p.numel()
            off += p.numel()
            if cur_n >= cap:
                self.buckets.append((off - cur_n, off, cur))
                cur, cur_n = [], 0
        if cur:
            self.buckets.append((off - cur_n, off, cur))
        for bi, (_, _, ps) in enumerate(self.buckets):
            for p in ps:
                self._where[id(p)] = bi
        self._pending = [len(ps) for _, _, ps in self.buckets]
        self._works, self._events = [], []
        self.cuda = dev.type == "cuda"
        self.stream = torch.cuda.Stream(device=dev) if self.cuda else None
        self.comm_ms = self.exposed_ms = 0.0
        self._hooks = [p.register_post_accumulate_grad_hook(self._ready) for p in self.params]

    @property
    def nbytes(self):
        return self.numel * 4

    def zero_grad(self):
        """Zero the flat buffer in place (``p.grad`` stays the bucket view) and re-arm the bucket counters."""
        self.flat.zero_()
        self._pending = [len(ps) for _, _, ps in self.buckets]
        self._works, self._events = [], []

    def _launch(self, bi):
        lo, hi, _ = self.buckets[bi]
        buf = self.flat[lo:hi]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())          # the gradients of this bucket are complete
            with torch.cuda.stream(self.stream):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                e1.record(self.stream)
            self._events.append((e0, e1))
        else:
            self._works.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _ready(self, p):
        if p.grad is not None and p.grad.data_ptr() != self._view_ptr(p):
            self._view(p).copy_(p.grad)                                     # somebody re-bound p.grad: fold it back into the bucket
            p.grad = self._view(p)
        bi = self._where[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _view(self, p):
        bi = self._where[id(p)]
        lo, _, ps = self.buckets[bi]
        off = lo
        for q in ps:
            if q is p:
                return self.flat[off:off + p.numel()].view_as(p)
            off += q.numel()
        raise KeyError

    def _view_ptr(self, p):
        return self._view(p).data_ptr()

    def finish(self):
        """Launch the buckets that never filled (parameters unused in this step contribute their zeros), wait for the side
        stream, average.  Call after ``loss.backward()`` and before the optimizer step."""
        for bi, n in enumerate(self._pending):
            if n > 0:
                self._pending[bi] = 0
                self._launch(bi)
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        if self.cuda:
            main = torch.cuda.current_stream()
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record(main)
            main.wait_stream(self.stream)
            w1.record(main)
            if world > 1:
                self.flat.div_(world)
            self._wait = (w0, w1)
        else:
            for w in self._works:
                w.wait()
            if world > 1:
                self.flat.div_(world)

    def stats(self):
        """(comm_ms, exposed_ms) of the last finished step; synchronises."""
        if not self.cuda:
            return 0.0, 0.0
        torch.cuda.synchronize()
        self.comm_ms = sum(a.elapsed_time(b) for a, b in self._events)
        self.exposed_ms = self._wait[0].elapsed_time(self._wait[1]) if hasattr(self, "_wait") else 0.0
        return self.comm_ms, self.exposed_ms
