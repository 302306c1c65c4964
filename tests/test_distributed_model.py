"""Two ranks, the REAL Voxel-DETR: the gradients after GradAverager.finish() equal the mean of the two single-process
gradients (same weights; each rank's scenes; BatchNorm statistics per rank, as the reference does not sync them —
sync_bn False, broadcast_buffers False, efg/engine/trainer.py:193-198; loss normalised by the world-mean number of
boxes, VD/losses.py:121-125).  SURVEY.md §8e's DDP-equivalence check.

CPU variant: oracle backend over gloo.  GPU variant: the CUDA kernels; NCCL when two GPUs are visible, otherwise both
ranks share cuda:0 over gloo (NCCL refuses two ranks on one device) — the bucket/hook/async logic is the same.
"""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(device, use_cuda):
    import model_cases as mc
    from efg_b200.detectors.voxel_detr import VoxelDETR

    cfg = mc.make_config("voxel_detr", device)
    if use_cuda:
        model = VoxelDETR(cfg)
    else:
        from oracle.backend_cpu import cpu_backend

        model = VoxelDETR(cfg, backend=cpu_backend())
    model.load_state_dict(mc.fill_state_dict(model.state_dict()))
    return cfg, model.train()


def _batch_for(rank, cfg):
    import model_cases as mc

    scenes = mc.make_scenes("voxel_detr")   # two scenes: rank r trains on scene r
    return mc.make_batch(scenes[rank:rank + 1], cfg.dataset)


def _worker(rank, world, port, out_dir, use_cuda, backend):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank if backend == "nccl" else 0))
    import torch.distributed as dist
    from efg_b200.parallel import GradAverager, init_distributed

    torch.set_num_threads(2)   # two workers share the host: oversubscribed OpenMP teams spin against each other
    if use_cuda:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    init_distributed(backend)
    device = ("cuda:%d" % (rank if backend == "nccl" else 0)) if use_cuda else "cpu"
    cfg, model = _build(device, use_cuda)
    avg = GradAverager(model, bucket_bytes=4 << 20)
    avg.broadcast_parameters()
    batch = _batch_for(rank, cfg)
    for it in range(2):   # step 0 learns the used-parameter set, step 1 reduces buckets from inside backward
        avg.zero_grad()
        losses = model([(dict(s), {"annotations": {k: v.copy() for k, v in i["annotations"].items()}}) for s, i in batch])
        total = sum(v for k, v in losses.items() if k.startswith("loss"))
        total.backward()
        launched_early = sum(1 for b in avg.buckets if b.launched)
        avg.finish()
        avg.hide_unused()
        if it == 0:   # undo the BN running-stat update so both steps see identical state (weights are not stepped)
            model.load_state_dict(__import__("model_cases").fill_state_dict(model.state_dict()))
    grads = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters() if p.grad is not None}
    torch.save({"grads": grads, "launched_early": launched_early, "num_buckets": len(avg.buckets),
                "loss": float(total.detach())}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def _single_process_expected(use_cuda, mean_boxes):
    """(g_rank0 + g_rank1) / 2 with the loss normaliser fixed to the world mean, computed in ONE process."""
    from efg_b200.detectors.voxel_detr.losses import Det3DLoss

    device = "cuda:0" if use_cuda else "cpu"
    sums = None
    threads = torch.get_num_threads()
    torch.set_num_threads(2)   # as in the workers: fp32 reductions depend on the thread count
    orig = Det3DLoss.__dict__["normaliser"]   # the staticmethod object itself (class attribute access would unwrap it)
    Det3DLoss.normaliser = staticmethod(lambda targets, dev: max(mean_boxes, 1.0))
    try:
        for rank in range(2):
            cfg, model = _build(device, use_cuda)
            losses = model(_batch_for(rank, cfg))
            sum(v for k, v in losses.items() if k.startswith("loss")).backward()
            g = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
            sums = g if sums is None else {n: sums[n] + g[n] for n in g}
    finally:
        Det3DLoss.normaliser = orig
        torch.set_num_threads(threads)
    return {n: v / 2 for n, v in sums.items()}


def _run(tmp_path, use_cuda, backend):
    import model_cases as mc

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), use_cuda, backend), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert set(r0["grads"]) == set(r1["grads"])
    for n in r0["grads"]:
        assert torch.equal(r0["grads"][n], r1["grads"][n]), n      # both ranks hold the same averaged gradient
    assert r0["launched_early"] >= 1 and r0["num_buckets"] >= 2     # overlap: buckets went out during backward
    scenes = mc.make_scenes("voxel_detr")
    mean_boxes = sum(len(a["labels"]) for _, a in scenes) / 2.0
    exp = _single_process_expected(use_cuda, mean_boxes)
    assert set(exp) == set(r0["grads"])
    # two runs of the same scene differ by summation order (thread count on the CPU, atomics and bf16x3 rounding on CUDA)
    # and whole-model gradients amplify that (DESIGN.md §5 note 2)
    worst, rels = 0.0, []
    for n, e in exp.items():
        if use_cuda:
            rels.append(((r0["grads"][n] - e).norm() / e.norm().clamp_min(1e-6)).item())
            continue
        scale = max(e.abs().max().item(), 1e-6)
        worst = max(worst, (r0["grads"][n] - e).abs().max().item() / scale)
        assert (r0["grads"][n] - e).abs().max().item() <= 1e-3 * scale + 1e-7, (n, worst)
    if use_cuda:
        # Two CUDA runs of the same scene are not bit-identical (atomics, split rounding), and at random initialisation a
        # flipped top-k proposal or Hungarian pair changes individual gradients by tens of per cent (measured: up to 29 %
        # in norm on single tensors; the same CPU graph in fp32 vs fp64 differs by ~7 %).  The CUDA variant
        # therefore checks the device plumbing statistically; the CPU variant above holds the arithmetic to 1e-3.
        rels = sorted(rels)
        # measured: median 5.8 %, i.e. the documented noise floor of this graph
        assert rels[len(rels) // 2] <= 0.15, rels[len(rels) // 2]
        assert rels[int(len(rels) * 0.9)] <= 0.5, rels[int(len(rels) * 0.9)]


def test_two_rank_voxel_detr_gradients_equal_mean_of_single_process_cpu(tmp_path):
    _run(tmp_path, use_cuda=False, backend="gloo")


@pytest.mark.gpu
def test_two_rank_voxel_detr_gradients_equal_mean_of_single_process_gpu(tmp_path):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    _run(tmp_path, use_cuda=True, backend=backend)
