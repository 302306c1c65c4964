"""GPU parity, whole model: Voxel-DETR on the CUDA backend vs the SAME module graph on the CPU
oracle backend, identical weights and synthetic scene.  north_star tolerance: logits / boxes within
1e-3 (fp32)."""
import numpy as np
import pytest
import torch

from efg_b200.detectors.voxel_detr import VoxelDETR
from oracle.backend_cpu import cpu_backend, voxelized_sample
from test_model_cpu import small_batch, small_config

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(autouse=True)
def _fp32_library_math():
    """The oracle is plain fp32; torch's cuDNN convolutions default to TF32 on the GPU (as they would
    for the reference too).  Parity is checked with the library layers in fp32."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


class _ReplayMatcher:
    """At random initialisation all queries are nearly identical, so the Hungarian cost matrix is
    nearly degenerate and 1e-6 differences flip assignments.  To compare kernels rather than
    tie-breaking, the GPU run replays the assignments of the CPU run."""

    def __init__(self):
        self.log = []
        self.replay = None

    def install(self, model, record):
        from efg_b200.detectors.voxel_detr.matcher import HungarianMatcher3d

        orig = HungarianMatcher3d.solve
        me = self

        def solve(mats):
            if record:
                out = orig(mats)
                me.log.append(out)
                return out
            return me.replay.pop(0)

        for head in (model.transformer.proposal_head, model.transformer.decoder.detection_head):
            head.losses.matcher.solve = solve


def _models(num_queries=40):
    torch.manual_seed(0)
    cpu = VoxelDETR(small_config("cpu", num_queries), backend=cpu_backend())
    gpu = VoxelDETR(small_config("cuda", num_queries))
    gpu.load_state_dict(cpu.state_dict())
    return cpu, gpu


def test_voxel_detr_eval_parity_points_in_voxels_out():
    cpu, gpu = _models()
    cpu.eval()
    gpu.eval()
    scenes = small_batch(2, 6000, seed=3)
    cfg = cpu.config
    batch_cpu = [(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in scenes]
    batch_gpu = [({"points": p}, {"annotations": a}) for p, a in scenes]  # voxelized on the GPU
    with torch.no_grad():
        fc, _ = cpu.extract(batch_cpu)
        fg, _ = gpu.extract(batch_gpu)
        assert (fg[0].cpu() - fc[0]).abs().max().item() < TOL
        hc = cpu.transformer(fc, [cpu.backbone.position_encoding(fc[0])])
        hg = gpu.transformer(fg, [gpu.backbone.position_encoding(fg[0])])
    # encoder memory
    assert (hg[3].cpu() - hc[3]).abs().max().item() < TOL
    # proposals are a top-k with sorted=False: compare as sets, then align the decoder outputs by index
    ic, ig = hc[5].squeeze(-1), hg[5].squeeze(-1).cpu()
    for b in range(ic.shape[0]):
        oc, og = torch.argsort(ic[b]), torch.argsort(ig[b])
        assert torch.equal(ic[b][oc], ig[b][og])
        assert (hg[0][:, b].cpu()[:, og] - hc[0][:, b][:, oc]).abs().max().item() < TOL      # hidden states
        assert (hg[2][:, b].cpu()[:, og] - hc[2][:, b][:, oc]).abs().max().item() < TOL      # refined boxes
        head_c, head_g = cpu.transformer.decoder.detection_head, gpu.transformer.decoder.detection_head
        lc, bc = head_c(hc[0][-1, b][oc], hc[2][-2, b][oc], 1)
        lg, bg = head_g(hg[0][-1, b][og.cuda()], hg[2][-2, b][og.cuda()], 1)
        assert (lg.cpu() - lc).abs().max().item() < TOL and (bg.cpu() - bc).abs().max().item() < TOL


def test_voxel_detr_train_step_parity():
    """Losses must agree to 1e-4.  Gradients of this graph at random initialisation are
    ill-conditioned: the SAME CPU oracle model evaluated in fp32 and in fp64 differs by ~7 % (median)
    in its gradients while its loss agrees to 7 digits (measured, DESIGN.md "parity notes").  Gradient
    parity of the kernels themselves is asserted where conditioning is benign (test_gpu_spconv.py,
    test_gpu_box_attn.py); here the CUDA path has to be as close to the CPU oracle as the oracle's own
    torch graph is when it merely runs on another device (the noise floor)."""
    cpu, gpu = _models()
    torch.manual_seed(0)
    gpo = VoxelDETR(small_config("cuda", 40), backend=cpu_backend())  # oracle ops (plain torch) on the GPU
    gpo.load_state_dict(cpu.state_dict())
    for m in (cpu, gpu, gpo):
        m.train()
    scenes = small_batch(2, 6000, seed=11)
    cfg = cpu.config
    rm = _ReplayMatcher()
    rm.install(cpu, record=True)
    lc = cpu([(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in scenes])
    rm.replay = list(rm.log)
    rm.install(gpu, record=False)
    lg = gpu([({"points": p}, {"annotations": a}) for p, a in scenes])
    rm.replay = list(rm.log)
    rm.install(gpo, record=False)
    lo = gpo([({"points": p}, {"annotations": a}) for p, a in scenes])
    assert set(lc) == set(lg)
    for k in lc:
        if k.startswith("loss"):
            assert abs(float(lg[k]) - float(lc[k])) < 1e-4 * max(1.0, abs(float(lc[k]))), (k, float(lg[k]), float(lc[k]))
    for losses in (lc, lg, lo):
        sum(v for k, v in losses.items() if k.startswith("loss")).backward()
    pc, pg, po = dict(cpu.named_parameters()), dict(gpu.named_parameters()), dict(gpo.named_parameters())
    checked, loose = 0, []
    for name, ref in pc.items():
        gc, gg, go = ref.grad, pg[name].grad, po[name].grad
        assert (gc is None) == (gg is None), name
        if gc is None:
            continue
        scale = max(float(gc.abs().max()), 1e-12)
        err = float((gg.cpu() - gc).abs().max()) / scale
        floor = float((go.cpu() - gc).abs().max()) / scale
        # the default bf16x3 tensor-core split rounds ~10x coarser per operation than fp32 (2.5e-5 vs 2.6e-6, both far
        # inside the 1e-3 bar), so the amplified gradient deviation may exceed the fp32 noise floor by a small factor.
        # The backward of box attention accumulates with atomics, so the factor itself moves from run to run: nearly
        # every parameter has to stay within 6x the floor, none may leave 12x.
        assert err < max(12.0 * floor, 1e-2), (name, err, floor)
        if not err < max(6.0 * floor, 5e-3):
            loose.append((name, err, floor))
        checked += 1
    assert checked > 150
    assert len(loose) <= checked // 20, loose


def test_conquer_train_losses_parity():
    """ConQueR on the CUDA backend vs the CPU oracle backend with identical denoising noise: every loss term
    (matching, denoising, query-contrast) within 1e-3 relative."""
    from efg_b200.config import conquer_config
    from efg_b200.detectors.conquer import ConQueR
    from efg_b200.detectors.conquer.cdn import draw_noise
    from test_model_cpu import SMALL

    def cfg(device):
        return conquer_config(dataset={"pc_range": SMALL.pc_range, "voxel_size": SMALL.voxel_size, "max_voxel_num": 20000},
                              model={"device": device, "transformer": {"num_queries": 40, "enc_layers": 1, "dec_layers": 2}})

    torch.manual_seed(0)
    cpu = ConQueR(cfg("cpu"), backend=cpu_backend()).train()
    gpu = ConQueR(cfg("cuda")).train()
    gpu.load_state_dict(cpu.state_dict())
    scenes = small_batch(2, 6000, seed=21)
    total_gt = sum(len(a["labels"]) for _, a in scenes)
    noise = draw_noise(2 * 3 * total_gt, 3, torch.device("cpu"), generator=torch.Generator().manual_seed(5))
    cpu.cdn_noise = gpu.cdn_noise = noise
    rm = _ReplayMatcher()
    rm.install(cpu, record=True)
    lc = cpu([(voxelized_sample(p, cpu.config.dataset), {"annotations": a}) for p, a in scenes])
    rm.replay = list(rm.log)
    rm.install(gpu, record=False)
    lg = gpu([({"points": p}, {"annotations": a}) for p, a in scenes])
    assert set(lc) == set(lg) and any(k.startswith("loss_contrastive") for k in lc) and "loss_ce_dn" in lc
    for k in lc:
        if k.startswith("loss"):
            assert abs(float(lg[k]) - float(lc[k])) < 1e-3 * max(1.0, abs(float(lc[k]))), (k, float(lg[k]), float(lc[k]))
    sum(v for k, v in lg.items() if k.startswith("loss")).backward()
    assert torch.isfinite(gpu.projector[0].weight.grad).all()


@pytest.mark.parametrize("family", ["voxel_detr", "conquer"])
def test_device_matching_equals_host_scipy_matching(family):
    """The same GPU model, same weights, same batch: Hungarian assignments solved on the device (csrc/lsa.cu) vs on
    the host with scipy (the reference's path).  Identical assignments -> identical losses."""
    from efg_b200.config import conquer_config
    from efg_b200.detectors.conquer import ConQueR
    from efg_b200.detectors.conquer.cdn import draw_noise
    from test_model_cpu import SMALL

    torch.manual_seed(0)
    if family == "voxel_detr":
        model = VoxelDETR(small_config("cuda", 40)).train()
    else:
        model = ConQueR(conquer_config(dataset={"pc_range": SMALL.pc_range, "voxel_size": SMALL.voxel_size, "max_voxel_num": 20000},
                                       model={"device": "cuda", "transformer": {"num_queries": 40, "enc_layers": 1,
                                                                                "dec_layers": 2}})).train()
    scenes = small_batch(2, 6000, seed=41)
    if family == "conquer":
        total_gt = sum(len(a["labels"]) for _, a in scenes)
        model.cdn_noise = draw_noise(2 * 3 * total_gt, 3, torch.device("cpu"), generator=torch.Generator().manual_seed(5))
    # non-degenerate predictions (at initialisation all queries are nearly identical and the optimum is not unique)
    with torch.no_grad():
        for p in model.transformer.decoder.detection_head.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    batch = [({"points": p}, {"annotations": a}) for p, a in scenes]
    out = {}
    import copy
    gt_state = copy.deepcopy(model.transformer.decoder_gt.state_dict()) if family == "conquer" else None
    for mode in (True, False):
        model.device_matching = mode
        if gt_state is not None:  # the EMA decoder moves at every training forward
            model.transformer.decoder_gt.load_state_dict(gt_state)
        torch.manual_seed(1)
        out[mode] = {k: float(v) for k, v in model(batch).items()}
    assert set(out[True]) == set(out[False])
    for k in out[True]:
        assert abs(out[True][k] - out[False][k]) <= 1e-5 * max(1.0, abs(out[False][k])), (k, out[True][k], out[False][k])


def test_centerpoint_train_losses_parity():
    """CenterPoint (SpMiddleResNetFHD with biased SubM blocks, padding [0,1,1] stage, un-padded z-collapse)
    on the CUDA backend vs the CPU oracle backend: losses within 1e-3 relative, BEV features within 1e-3."""
    from efg_b200.config import centerpoint_config
    from efg_b200.detectors.centerpoint import VoxelNet
    from test_model_cpu import SMALL

    def cfg(device):
        return centerpoint_config(dataset={"pc_range": SMALL.pc_range, "voxel_size": SMALL.voxel_size,
                                           "max_voxel_num": 20000}, model={"device": device})

    torch.manual_seed(0)
    cpu = VoxelNet(cfg("cpu"), backend=cpu_backend()).train()
    gpu = VoxelNet(cfg("cuda")).train()
    gpu.load_state_dict(cpu.state_dict())
    scenes = small_batch(2, 6000, seed=31)
    batch_cpu = [(voxelized_sample(p, cpu.config.dataset), {"annotations": a}) for p, a in scenes]
    batch_gpu = [({"points": p}, {"annotations": a}) for p, a in scenes]
    lc, lg = cpu(batch_cpu), gpu(batch_gpu)
    for k in lc:
        assert abs(float(lg[k]) - float(lc[k])) < 1e-3 * max(1.0, abs(float(lc[k]))), (k, float(lg[k]), float(lc[k]))
    lc["0_loss"].backward()
    lg["0_loss"].backward()
    # the head's last layers are well conditioned: gradients must agree
    gc = dict(cpu.named_parameters())["center_head.tasks.0.hm.3.weight"].grad
    gg = dict(gpu.named_parameters())["center_head.tasks.0.hm.3.weight"].grad
    assert (gg.cpu() - gc).abs().max().item() < 1e-3 * max(1.0, gc.abs().max().item())


def test_prepared_geometry_gives_the_same_step():
    """VoxelDETR.prepare(): voxelizer + strided rulebooks built ahead on another stream; forward(prepared=...) must give
    the losses and gradients of the in-line path, and must not build any strided rulebook itself."""
    from efg_b200 import ops

    _, gpu = _models()
    gpu.train()
    scenes = small_batch(2, 6000, seed=5)
    inputs = [({"points": torch.from_numpy(p).cuda()}, {"annotations": a}) for p, a in scenes]
    torch.manual_seed(1)
    plain = gpu(inputs)
    sum(v for k, v in plain.items() if k.startswith("loss")).backward()
    g_plain = {n: p.grad.clone() for n, p in gpu.named_parameters() if p.grad is not None}
    gpu.zero_grad(set_to_none=True)

    side = torch.cuda.Stream(priority=-1)
    prepared = gpu.prepare(inputs, side)
    assert prepared is not None and len(prepared["indice_dict"]) >= 3
    calls = []
    real = ops.sparse_rulebook
    ops.sparse_rulebook = lambda *a, **k: calls.append(1) or real(*a, **k)
    try:
        torch.manual_seed(1)
        ahead = gpu(inputs, prepared=prepared)
    finally:
        ops.sparse_rulebook = real
    assert not calls, "the prepared step built %d strided rulebooks in line" % len(calls)
    assert set(ahead) == set(plain)
    for k in plain:
        assert torch.allclose(ahead[k], plain[k], rtol=1e-5, atol=1e-6), (k, float(ahead[k]), float(plain[k]))
    sum(v for k, v in ahead.items() if k.startswith("loss")).backward()
    for n, p in gpu.named_parameters():
        if n in g_plain:
            scale = float(g_plain[n].abs().max()) + 1e-12
            assert float((p.grad - g_plain[n]).abs().max()) / scale < 5e-3, n   # atomics in the attention backward
