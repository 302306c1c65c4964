"""CPU, world_size 2 over gloo: the data-parallel plumbing (parameter broadcast, gradient buckets reduced from inside
backward, unused parameters, the asynchronous num_boxes all-reduce of the loss) reproduces single-process gradients."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from efg_b200.detectors.voxel_detr.losses import Det3DLoss
    from efg_b200.parallel import GradAverager, init_distributed

    r, _, w = init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)  # different initial weights per rank: broadcast must fix that
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    unused = torch.nn.Linear(3, 3)  # never used: has no gradient on any rank (like the pruned FPN levels)
    holder = torch.nn.ModuleList([model, unused])
    avg = GradAverager(holder, bucket_bytes=256)  # tiny buckets -> several all-reduces
    avg.broadcast_parameters()
    torch.manual_seed(7)
    data = torch.randn(3, 2, 6, 8)
    nbytes = 0
    for it in range(3):  # step 0 learns which parameters receive gradients; steps 1-2 launch the buckets from the hooks
        avg.zero_grad()
        loss = model(data[it, rank]).square().mean()
        loss.backward()
        if it > 0:
            assert any(b.launched for b in avg.buckets), "no bucket was reduced from inside backward"
        nbytes = avg.finish()
        avg.hide_unused()
    assert nbytes == sum(p.numel() * 4 for p in holder.parameters())  # fixed-size buckets: independent of the data
    assert all(p.grad is None for p in unused.parameters())
    # loss normaliser: mean number of boxes per rank, all-reduced (VD/losses.py:121-125)
    targets = [{"labels": torch.zeros(3 + 4 * rank, dtype=torch.long)}]
    nb = Det3DLoss.normaliser(targets, torch.device("cpu"))
    torch.save({"grads": [p.grad.clone() for p in model.parameters()], "params": [p.detach().clone() for p in model.parameters()],
                "num_boxes": nb}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average_matches_single_process(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)  # broadcast from rank 0
    for a, b in zip(r0["grads"], r1["grads"]):
        assert torch.equal(a, b)  # both ranks hold the averaged gradient
    assert r0["num_boxes"] == r1["num_boxes"] == (3 + 7) / 2
    # single-process reference: mean of the two per-rank losses
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    with torch.no_grad():
        for p, q in zip(model.parameters(), r0["params"]):
            p.copy_(q)
    torch.manual_seed(7)
    data = torch.randn(3, 2, 6, 8)[2]  # the last of the three steps
    (0.5 * (model(data[0]).square().mean() + model(data[1]).square().mean())).backward()
    for p, g in zip(model.parameters(), r0["grads"]):
        assert torch.allclose(p.grad, g, atol=1e-6)


def test_grad_averager_single_process_drops_its_hooks_and_follows_late_parameters():
    """world size 1: after the first step the per-parameter hooks are gone (they only taught which parameters receive
    gradients), unused parameters keep grad = None towards the optimizer, and a parameter that starts to receive
    gradients later is used from then on."""
    import torch
    from efg_b200.parallel import GradAverager

    torch.manual_seed(0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)
            self.use_b = False

        def forward(self, x):
            y = self.a(x)
            return self.b(y) if self.use_b else y

    net = Net()
    av = GradAverager(net)
    x = torch.randn(3, 4)
    seen = []
    for step in range(4):
        if step == 2:
            net.use_b = True
        av.zero_grad()
        net(x).sum().backward()
        av.finish()
        av.hide_unused()
        seen.append([p.grad is not None for p in net.parameters()])
        # gradients are the bucket views (the optimizer and the all-reduce share them)
        for b in av.buckets:
            for p in b.params:
                if p.grad is not None:
                    assert p.grad.untyped_storage().data_ptr() == b.flat.untyped_storage().data_ptr()
    assert seen[0] == [True, True, False, False] and seen[1] == seen[0]
    assert seen[2] == [True, True, True, True] and seen[3] == seen[2]
    assert len(av._hooks) == 2          # only the two parameters that were unused in the first step are still watched
    ref = Net()
    ref.load_state_dict(net.state_dict())
    ref.use_b = True
    ref(x).sum().backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad)
