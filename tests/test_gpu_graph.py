"""GPU: the CUDA-graphed static section (FPN top-down -> transformer -> heads, forward AND backward replayed from graphs,
VoxelDETR.enable_static_graph) gives the same losses and gradients as the eager path, on the batch it was captured with
and on a different batch (different voxel counts, different ground truth: the dynamic parts stay eager)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


def _step(model, batch):
    for p in model.parameters():
        p.grad = None
    losses = model([({"points": torch.from_numpy(p).cuda()}, {"annotations": {k: v.copy() for k, v in a.items()}}) for p, a in batch])
    total = sum(v for k, v in losses.items() if k.startswith("loss"))
    total.backward()
    torch.cuda.synchronize()
    return ({k: float(v.detach()) for k, v in losses.items()},
            {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})


def test_static_graph_equals_eager():
    import model_cases as mc
    from efg_b200.data import SceneSpec, make_scene
    from efg_b200.detectors.voxel_detr import VoxelDETR

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # graph capture needs the job on a non-default stream (see bench.py / VoxelDETR.enable_static_graph)
    with torch.cuda.stream(torch.cuda.Stream()):
        cfg = mc.make_config("voxel_detr", "cuda")
        torch.manual_seed(0)
        model = VoxelDETR(cfg).train()
        model.load_state_dict(mc.fill_state_dict(model.state_dict()))
        for m in model.modules():   # frozen BN statistics: eager and graphed runs must see identical state
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 0.0
        spec = SceneSpec(pc_range=mc.SMALL_RANGE, voxel_size=mc.VOXEL)
        batches = [[make_scene(4000 + 500 * b, spec, seed=10 * b + i, num_objects=5 + b) for i in range(2)] for b in range(2)]
        eager = [_step(model, b) for b in batches]
        assert model.enable_static_graph([({"points": torch.from_numpy(p).cuda()}, {"annotations": a}) for p, a in batches[0]]), \
            model.static_graph_error
        graphed = [_step(model, b) for b in batches]
        graphed_again = _step(model, batches[0])   # replays do not carry state from the previous batch
    for (le, ge), (lg, gg) in zip(eager + [eager[0]], graphed + [graphed_again]):
        assert set(le) == set(lg)
        for k in le:
            assert abs(le[k] - lg[k]) <= 1e-4 * max(1.0, abs(le[k])), (k, le[k], lg[k])
        assert set(ge) == set(gg)
        rels = sorted(((ge[n] - gg[n]).norm() / ge[n].norm().clamp_min(1e-6)).item() for n in ge)
        # same kernels, same inputs: differences are atomics order only (see test_distributed_model.py for the noise floor)
        assert rels[len(rels) // 2] <= 0.05 and rels[int(len(rels) * 0.9)] <= 0.3, (rels[len(rels) // 2], rels[-1])


def test_centerpoint_static_graph_and_prepared_step_match_eager():
    """CenterPoint: neck + heads + losses in CUDA graphs, fed by prepare() from a side stream, give the eager losses and
    gradients."""
    from efg_b200.config import centerpoint_config
    from efg_b200.detectors.centerpoint import VoxelNet
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_model_cpu import SMALL, small_batch

    torch.manual_seed(0)
    cfg = centerpoint_config(dataset={"pc_range": SMALL.pc_range, "voxel_size": SMALL.voxel_size, "max_voxel_num": 20000},
                             model={"device": "cuda"})
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):   # captures need a non-default stream for the whole job
        model = VoxelNet(cfg).train()
        scenes = small_batch(2, 8000, seed=2)
        inputs = [({"points": torch.from_numpy(p).cuda()}, {"annotations": a}) for p, a in scenes]

        def run(prepared=None):
            model.zero_grad(set_to_none=True)
            losses = model(inputs) if prepared is None else model(inputs, prepared=prepared)
            sum(v for k, v in losses.items() if k.endswith("_loss") and k.count("_") == 1).backward()
            return {k: float(v) for k, v in losses.items()}, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

        for m in model.modules():   # identical batch statistics on every call: freeze the running averages
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 0.0
        for _ in range(2):
            eager_l, eager_g = run()
        assert model.enable_static_graph(inputs), model.static_graph_error
        prep = model.prepare(inputs, torch.cuda.Stream(priority=-1))
        assert prep is not None and prep["targets"] is not None
        graph_l, graph_g = run(prep)
        torch.cuda.synchronize()
    assert set(graph_l) == set(eager_l)
    for k in eager_l:
        assert abs(graph_l[k] - eager_l[k]) <= 1e-4 * max(1.0, abs(eager_l[k])), (k, graph_l[k], eager_l[k])
    assert set(graph_g) == set(eager_g)
    for n in eager_g:
        scale = float(eager_g[n].abs().max())
        if scale < 1e-3:
            continue   # conv biases in front of a BatchNorm: the true gradient is zero, what is left is rounding noise
        assert float((graph_g[n] - eager_g[n]).abs().max()) / scale < 2e-3, n


def test_conquer_encoder_graph_equals_eager():
    """ConQueR: the encoder section replayed from CUDA graphs (decoders and losses eager) gives the eager losses and
    gradients with identical denoising noise, on the captured batch and on another one."""
    import model_cases as mc
    from efg_b200.data import SceneSpec, make_scene
    from efg_b200.detectors.conquer import ConQueR

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.cuda.stream(torch.cuda.Stream()):
        cfg = mc.make_config("conquer", "cuda")
        torch.manual_seed(0)
        model = ConQueR(cfg).train()
        model.load_state_dict(mc.fill_state_dict(model.state_dict()))
        for m in model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 0.0
        model.transformer.m = 1.0   # frozen momentum decoder: every run sees the same state
        spec = SceneSpec(pc_range=mc.SMALL_RANGE, voxel_size=mc.VOXEL)
        batches = [[make_scene(4000 + 500 * b, spec, seed=10 * b + i, num_objects=5 + b) for i in range(2)] for b in range(2)]

        def run(b):
            torch.manual_seed(7)   # the denoising noise is drawn from the global generator
            return _step(model, b)

        eager = [run(b) for b in batches]
        assert model.enable_static_graph([({"points": torch.from_numpy(p).cuda()}, {"annotations": a}) for p, a in batches[0]]), \
            model.static_graph_error
        graphed = [run(b) for b in batches]
    for (le, ge), (lg, gg) in zip(eager, graphed):
        assert set(le) == set(lg)
        for k in le:
            assert abs(le[k] - lg[k]) <= 1e-4 * max(1.0, abs(le[k])), (k, le[k], lg[k])
        assert set(ge) == set(gg)
        rels = sorted(((ge[n] - gg[n]).norm() / ge[n].norm().clamp_min(1e-6)).item() for n in ge)
        assert rels[len(rels) // 2] <= 0.05 and rels[int(len(rels) * 0.9)] <= 0.3, (rels[len(rels) // 2], rels[-1])


def test_graphed_optimizer_step_matches_eager_and_bumps_versions():
    """parallel.GraphedOptimizerStep: AdamW replayed from a CUDA graph updates the parameters like the eager step, over
    several steps with changing gradients, and bumps Tensor._version (the packed-weight cache is keyed on it)."""
    from efg_b200.parallel import GradAverager, GraphedOptimizerStep

    torch.manual_seed(0)
    with torch.cuda.stream(torch.cuda.Stream()):
        def make():
            torch.manual_seed(3)
            return torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.ReLU(), torch.nn.Linear(64, 8)).cuda()
        a, b = make(), make()
        opts = [torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.01, betas=(0.9, 0.99), eps=1e-9, fused=True, capturable=True)
                for m in (a, b)]
        avs = [GradAverager(m) for m in (a, b)]
        xs = [torch.randn(16, 32, device="cuda") for _ in range(6)]

        def backward(m, av, x):
            av.zero_grad()
            m(x).square().mean().backward()
            av.finish()
            av.hide_unused()

        for m, av, o in zip((a, b), avs, opts):      # one eager step each: optimizer state exists, buckets in place
            backward(m, av, xs[0])
            o.step()
        backward(b, avs[1], xs[1])                   # gradients in place for the capture (the capture itself applies nothing)
        graphed = GraphedOptimizerStep(opts[1])
        for x in xs[1:]:
            backward(a, avs[0], x)
            opts[0].step()
            backward(b, avs[1], x)
            v0 = [p._version for p in b.parameters()]
            graphed.step()
            assert all(p._version > v for p, v in zip(b.parameters(), v0))
        torch.cuda.synchronize()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-7), float((pa - pb).abs().max())
