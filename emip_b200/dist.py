"""Multi-GPU plumbing for the hot path: batch sharding and the training gradient all-reduce.

The path shards by frame pairs / clips with no data-path collective.  The one collective of the
reference is DDP's gradient all-reduce (train.py:279); ``FlatGradAllReduce`` is the explicit equivalent for the
learnable prompt / decoder parameters: gradients are packed into ONE flat fp32 buffer and reduced with ONE
``all_reduce`` (NCCL over NVLink on the GPU box, gloo in the CPU tests), which replaces DDP's
``find_unused_parameters=True`` graph walk -- parameters without a gradient contribute zeros, exactly as DDP does.
"""
import torch
import torch.distributed as dist


def shard_batch(n_items, world_size, rank):
    """Contiguous [start, stop) slice of ``n_items`` frame pairs / clips owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FlatGradAllReduce:
    """Average the gradients of ``params`` across the process group through one flat buffer."""

    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.numel = sum(p.numel() for p in self.params)
        if not self.params:
            raise ValueError("no trainable parameters")
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @property
    def nbytes(self):
        return self.numel * 4

    def __call__(self):
        """Pack -> all_reduce(SUM) -> divide by world size -> unpack into ``p.grad``."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()                      # unused parameter on this rank (the reference needs find_unused_parameters)
            else:
                v.copy_(p.grad)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(dist.get_world_size(self.group))
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
