"""Data parallelism over scenes: one process per GPU, gradients averaged with NCCL all-reduce.

The reference's only parallelism is DistributedDataParallel (efg/engine/trainer.py:191-198, find_unused_parameters
from the playground configs).  The path shards over scenes with no data-path exchange, so the collectives are the
gradient all-reduce and the 1-float ``num_boxes`` all-reduce of the loss (VD/losses.py:121-125).

``GradAverager`` gives DDP's semantics without its wrapper:
  * the trainable parameters are laid out ONCE into a few flat buckets, in reverse registration order (roughly the
    order backward produces them), and every ``p.grad`` is a view into its bucket — no flatten / copy-back kernels;
  * a post-accumulate-grad hook per parameter counts arrivals; the all-reduce of a bucket is launched from inside
    backward as soon as the last expected gradient of that bucket has been accumulated, so NCCL overlaps the rest of
    backward (round 1 reduced serially after backward: a flat +3 ms per step for N >= 2);
  * which parameters are "expected" is learned from the first step: parameters that received no gradient there (the
    pruned FPN levels; DDP's unused parameters) stop being waited for.  A parameter that misses a later step (a
    data-dependent branch, e.g. a rank without ground truth) only delays its bucket until ``finish()``; its slot holds
    zeros, i.e. it contributes nothing to the mean — the bucket sizes never depend on the data, so ranks cannot
    disagree about the collective (the failure mode of reducing "the gradients that exist");
  * parameters that received no gradient on this rank in the first step keep ``grad = None`` towards the optimizer
    (``hide_unused``), as under DDP, so AdamW applies neither weight decay nor momentum to them.
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


class _Bucket:
    __slots__ = ("flat", "params", "expected", "arrived", "handle", "launched")

    def __init__(self, flat, params):
        self.flat, self.params = flat, params
        self.expected = len(params)
        self.arrived = 0
        self.handle = None
        self.launched = False


class GradAverager:
    def __init__(self, module, bucket_bytes=16 << 20, overlap=True):
        self.module = module
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = overlap
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.buckets = []
        self._bucket_of = {}
        self._fired = set()
        self._used = None          # ids of the parameters that received a gradient in the first step
        self._steps = 0
        self._avg_op = None
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self._make_bucket(cur)
                cur, size = [], 0
        if cur:
            self._make_bucket(cur)
        # the hooks also run in a single process: they are how "received a gradient" is known (hide_unused)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._unused = None        # parameters hide_unused() detaches (known after the first step)

    def _make_bucket(self, params):
        p0 = params[0]
        flat = torch.zeros(sum(p.numel() for p in params), dtype=p0.dtype, device=p0.device)
        b = _Bucket(flat, list(params))
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            self._bucket_of[id(p)] = b
        self.buckets.append(b)

    # ---- per step ---------------------------------------------------------------------------------
    @torch.no_grad()
    def zero_grad(self):
        """One memset per bucket; re-attaches the gradient views (use instead of optimizer.zero_grad)."""
        for b in self.buckets:
            b.flat.zero_()
            b.arrived, b.handle, b.launched = 0, None, False
        if self._unused is not None:
            # steady state: only the parameters hide_unused() detached need their view back (nothing else replaces a
            # gradient: the optimizer and autograd write through the views)
            for p, view in self._unused:
                p.grad = view
        else:
            for b in self.buckets:
                off = 0
                for p in b.params:
                    if p.grad is None or p.grad.data_ptr() != b.flat.data_ptr() + off * b.flat.element_size():
                        p.grad = b.flat[off:off + p.numel()].view_as(p)
                    off += p.numel()
        self._fired.clear()

    def _on_grad(self, p):
        self._fired.add(id(p))
        b = self._bucket_of[id(p)]
        if self._used is not None and id(p) not in self._used:
            if self.world > 1 and self.overlap and b.launched:
                # the bucket is already being reduced: adding to it now would race with NCCL.  Same contract as DDP's
                # static graph: the set of parameters that receive gradients is fixed after the first step.
                raise RuntimeError("GradAverager(overlap=True): a parameter received its first gradient after the first "
                                   "step, while its bucket was already in flight; construct with overlap=False for "
                                   "models whose set of used parameters changes between steps")
            self._used.add(id(p))   # expected from the next step on; this step it is reduced with its bucket at finish()
            return
        b.arrived += 1
        if self._used is not None and self.world > 1 and self.overlap and b.arrived == b.expected and not b.launched:
            self._launch(b)

    def _on_late_grad(self, p):
        self._used.add(id(p))
        self._unused = [(q, v) for q, v in self._unused if q is not p]

    def _launch(self, b):
        b.launched = True
        if self._avg_op is None:
            self._avg_op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        b.handle = dist.all_reduce(b.flat, op=self._avg_op, async_op=True)

    @torch.no_grad()
    def finish(self):
        """After backward: launch what has not been launched, wait (stream-ordered on CUDA), return the bytes reduced."""
        if self.world == 1:
            self._after_step()
            return 0
        for b in self.buckets:
            if not b.launched:
                self._launch(b)
        total = 0
        for b in self.buckets:
            b.handle.wait()
            if self._avg_op == dist.ReduceOp.SUM:
                b.flat.div_(self.world)
            total += b.flat.numel() * b.flat.element_size()
        self._after_step()
        return total

    average_gradients = finish  # round-1 name

    def _after_step(self):
        if self._used is None:
            self._used = set(self._fired)
            for b in self.buckets:
                b.expected = sum(1 for p in b.params if id(p) in self._used)
            if self.world == 1:
                # a single process has nothing to launch from the hooks: they were only needed to learn which parameters
                # receive gradients (262 Python calls per backward otherwise)
                for h in self._hooks:
                    h.remove()
                self._hooks = []
                self._unused = []
                for b in self.buckets:
                    off = 0
                    for p in b.params:
                        if id(p) not in self._used:
                            self._unused.append((p, b.flat[off:off + p.numel()].view_as(p)))
                            # a parameter that starts to receive gradients later is used from then on
                            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_late_grad))
                        off += p.numel()
        self._steps += 1

    def hide_unused(self):
        """Set ``grad = None`` on the parameters that got no gradient in the first step (DDP leaves them None, so the
        optimizer skips them); zero_grad() re-attaches the views.  Call between finish() and optimizer.step()."""
        if self._used is None:
            return
        if self._unused is not None:
            for p, _ in self._unused:
                p.grad = None
            return
        for p in self.params:
            if id(p) not in self._used:
                p.grad = None

    @torch.no_grad()
    def broadcast_parameters(self, src=0):
        if self.world == 1:
            return
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            dist.broadcast(t.data, src)


class AsyncMean:
    """`sum over ranks / world` of one host float, requested early and consumed late: the all-reduce is issued
    asynchronously when the targets are known (start of forward) and waited for where the loss needs it, so ranks of a
    launch-bound step are not forced into lock-step in the middle of the loss (VD/losses.py:121-125 blocks there)."""

    def __init__(self, value, device):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.local = float(value)
        self.handle = None
        self.t = None
        if self.world > 1:
            self.t = torch.as_tensor([self.local], dtype=torch.float, device=device)
            self.handle = dist.all_reduce(self.t, async_op=True)

    def result(self, floor=1.0):
        if self.world == 1:
            return max(self.local, floor)
        self.handle.wait()
        if self.t.is_cuda:  # stay on the device: an .item() would drain the pipeline once per step on every rank
            return (self.t / self.world).clamp_(min=floor).reshape(())
        return max(float(self.t.item()) / self.world, floor)


class GraphedOptimizerStep:
    """``optimizer.step()`` replayed from a CUDA graph.  The multi-tensor optimizers of torch regroup their tensor lists in
    Python on every call (~3 ms for the 260 parameters of Voxel-DETR); in a training loop whose gradients live at fixed
    addresses — the bucket views of GradAverager — the step is the same kernels on the same pointers every time.

    Construct it after a backward pass, with every gradient that will ever exist in place (and the unused ones hidden:
    GradAverager.hide_unused), from an optimizer created with ``capturable=True`` (its step counters then live on the
    device).  Hyper-parameters held as Python numbers (lr, betas, weight decay) are baked in: re-capture after changing
    them, or keep lr in a tensor.  A replay updates the parameters in place without going through autograd's version
    counters, so they are bumped by hand: everything keyed on ``Tensor._version`` (the packed weight images of
    efg_b200.ops) sees the update.

    Not wired into bench.py: three A/B runs of the Voxel-DETR step showed the device-resident arm SLOWER with it (51–53 vs
    58–59 scenes/s, host enqueue 37 vs 32 ms) while the end-to-end arm was unchanged; the cause was not found within the
    round's GPU budget.  The class and its equality test (tests/test_gpu_graph.py) stay as a tool."""

    def __init__(self, optimizer):
        self.optimizer = optimizer
        self.params = [p for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
        if not self.params or not all(p.is_cuda for p in self.params):
            raise RuntimeError("GraphedOptimizerStep needs CUDA parameters with gradients in place")
        if not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise RuntimeError("GraphedOptimizerStep needs an optimizer created with capturable=True")
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            optimizer.step()

    def step(self):
        self.graph.replay()
        torch.autograd.graph.increment_version(self.params)
