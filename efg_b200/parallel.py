"""Data parallelism over scenes: one process per GPU, gradients averaged with NCCL all-reduce.

The reference's only parallelism is DistributedDataParallel (efg/engine/trainer.py:191-198).  The
path shards over scenes with no data-path exchange, so the only collective is the gradient
all-reduce (plus the 1-float ``num_boxes`` all-reduce inside the loss).  ``GradAverager`` flattens
the gradients that exist (parameters of pruned branches have none, on every rank alike — the
reference needs ``find_unused_parameters`` for the same parameters) into a few large buckets and
all-reduces them; ~73 MB over NVLink 5 is ~0.2 ms, so no overlap machinery is needed.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment. Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


class GradAverager:
    def __init__(self, module, bucket_bytes=64 << 20):
        self.module = module
        self.bucket_bytes = bucket_bytes
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    @torch.no_grad()
    def broadcast_parameters(self, src=0):
        if self.world == 1:
            return
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            dist.broadcast(t.data, src)

    @torch.no_grad()
    def average_gradients(self):
        """All-reduce (mean) every existing gradient, bucketed; returns the number of bytes reduced."""
        if self.world == 1:
            return 0
        grads = [p.grad for p in self.module.parameters() if p.grad is not None]
        buckets, bucket, size, total = [], [], 0, 0
        for g in grads:
            bucket.append(g)
            size += g.numel() * g.element_size()
            total += g.numel() * g.element_size()
            if size >= self.bucket_bytes:
                buckets.append(bucket)
                bucket, size = [], 0
        if bucket:
            buckets.append(bucket)
        handles = []
        for members in buckets:
            flat = torch.cat([g.reshape(-1) for g in members])  # one batched copy kernel
            handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, members))
        for h, flat, members in handles:
            h.wait()
            flat.div_(self.world)
            views = [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in members]), members)]
            torch._foreach_copy_(members, views)  # multi-tensor copy back: a handful of launches, not one per gradient
        return total
