"""BEV IoU of rotated boxes / rotated NMS.

Goldens (tests/golden/iou3d_*.npz) were produced by the REFERENCE's own iou3d_cpu.cpp, compiled from /root/reference
(tests/golden/make_golden_iou3d.py).  CPU: the oracle's C restatement must equal them bit for bit (IoU matrix and NMS
selections).  GPU: the kernels (csrc/iou3d.cu) through the efg._C-style entry points must give the same IoU values to
fp32 rounding (device sinf / cosf / atan2f differ from libm in the last ulp) and EXACTLY the same kept indices — every
golden records how far its decisive IoU values are from the thresholds (>= 5e-4), far outside that rounding."""
import glob
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDENS = sorted(glob.glob(os.path.join(HERE, "golden", "iou3d_*.npz")))
THRESHOLDS = (0.1, 0.5, 0.7)


def test_goldens_exist():
    assert len(GOLDENS) >= 5


@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p)[6:-4] for p in GOLDENS])
def test_oracle_matches_reference_golden(path):
    from oracle import iou3d

    g = np.load(path)
    assert np.array_equal(iou3d.boxes_iou_bev(g["boxes_a"], g["boxes_b"]), g["iou"])
    assert np.array_equal(iou3d.boxes_iou_bev(g["boxes_a"], g["boxes_a"]), g["self_iou"])
    for thr in THRESHOLDS:
        assert np.array_equal(iou3d.nms(g["boxes_a"], thr), g["keep_%g" % thr])


def test_oracle_properties():
    from oracle import iou3d

    rng = np.random.default_rng(3)
    b = np.concatenate([rng.uniform(-5, 5, (40, 3)), rng.uniform(0.5, 4, (40, 3)), rng.uniform(-3, 3, (40, 1))], 1).astype(np.float32)
    iou = iou3d.boxes_iou_bev(b, b)
    assert np.allclose(np.diag(iou), 1.0, atol=1e-4) and np.allclose(iou, iou.T, atol=1e-4) and iou.min() >= 0 and iou.max() <= 1 + 1e-4
    area = iou3d.boxes_iou_bev(b, b, overlap=True)
    assert np.allclose(np.diag(area), b[:, 3] * b[:, 4], rtol=1e-4)
    assert len(iou3d.nms(b[:0], 0.5)) == 0 and iou3d.nms(b[:1], 0.5).tolist() == [0]


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p)[6:-4] for p in GOLDENS])
def test_gpu_iou_and_nms_match_reference_golden(path):
    from efg_b200 import _C
    from efg_b200.operators import iou3d_nms

    g = np.load(path)
    a, b = torch.from_numpy(g["boxes_a"]).cuda(), torch.from_numpy(g["boxes_b"]).cuda()
    ans = torch.zeros(a.shape[0], b.shape[0], device="cuda")
    assert _C.boxes_iou_bev_gpu(a, b, ans) == 1
    assert np.abs(ans.cpu().numpy() - g["iou"]).max() < 2e-5
    assert np.abs(iou3d_nms.boxes_iou_bev(a, b).cpu().numpy() - g["iou"]).max() < 2e-5
    for thr in THRESHOLDS:
        assert float(g["margin_%g" % thr]) > 1e-4   # the decision is not within rounding of the threshold
        keep = torch.zeros(a.shape[0], dtype=torch.int64)
        n = _C.nms_gpu(a, keep, thr)
        assert keep[:n].tolist() == g["keep_%g" % thr].tolist()
        # the Python wrapper sorts by score first: descending scores = the golden's order
        scores = torch.arange(a.shape[0], 0, -1, device="cuda", dtype=torch.float32)
        sel, _ = iou3d_nms.nms_gpu(a, scores, thr)
        assert sel.cpu().tolist() == g["keep_%g" % thr].tolist()


@pytest.mark.gpu
def test_gpu_nms_large_and_edge_cases():
    """4096 boxes (CenterPoint's nms_pre_max_size): 64 mask blocks, the device scan vs the oracle; axis-aligned variant;
    empty and single-box inputs; CPU tensors raise as the reference's CHECK_INPUT does."""
    from efg_b200 import _C, ops
    from oracle import iou3d

    rng = np.random.default_rng(11)
    n = 4096
    b = np.concatenate([rng.uniform(-60, 60, (n, 2)), rng.uniform(-1, 1, (n, 1)), rng.uniform(0.6, 1.4, (n, 3)) * [4.5, 2.0, 1.6],
                        rng.uniform(-np.pi, np.pi, (n, 1))], 1).astype(np.float32)
    gb = torch.from_numpy(b).cuda()
    for normal in (False, True):
        keep, count = ops.nms_bev(gb, 0.2, normal=normal)
        exp = iou3d.nms(b, 0.2, normal=normal)
        got = keep[:int(count.item())].cpu().numpy()
        # decisions within fp32 rounding of the threshold may differ between libm and the device: compare and allow none
        # to be far off — on this seed the selections are identical
        assert np.array_equal(got, exp), (len(got), len(exp))
    keep = torch.zeros(0, dtype=torch.int64)
    assert _C.nms_gpu(gb[:0].contiguous(), keep, 0.5) == 0
    keep = torch.zeros(1, dtype=torch.int64)
    assert _C.nms_gpu(gb[:1].contiguous(), keep, 0.5) == 1 and keep.tolist() == [0]
    with pytest.raises(RuntimeError):
        _C.nms_gpu(torch.from_numpy(b), torch.zeros(n, dtype=torch.int64), 0.5)
    ov = torch.zeros(8, 8, device="cuda")
    _C.boxes_overlap_bev_gpu(gb[:8].contiguous(), gb[:8].contiguous(), ov)
    assert np.allclose(ov.cpu().numpy(), iou3d.boxes_iou_bev(b[:8], b[:8], overlap=True), atol=1e-4)


@pytest.mark.gpu
def test_centerpoint_eval_applies_rotated_nms():
    """CenterHead.decode on the CUDA path == the same head wired over the oracle NMS (duplicates are suppressed)."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import model_cases as mc
    from efg_b200.detectors.centerpoint import VoxelNet
    from oracle.backend_cpu import cpu_backend

    torch.backends.cudnn.allow_tf32 = False
    cfg_g, cfg_c = mc.make_config("centerpoint", "cuda"), mc.make_config("centerpoint", "cpu")
    for cfg in (cfg_g, cfg_c):
        cfg.model.post_process.score_threshold = 0.3
    gpu, cpu = VoxelNet(cfg_g).eval(), VoxelNet(cfg_c, backend=cpu_backend()).eval()
    sd = mc.fill_state_dict(cpu.state_dict())
    gpu.load_state_dict(sd)
    cpu.load_state_dict(sd)
    scenes = mc.make_scenes("centerpoint")[:1]
    with torch.no_grad():
        rg = gpu(mc.make_batch(scenes, cfg_g.dataset))[0]
        rc = cpu(mc.make_batch(scenes, cfg_c.dataset))[0]
    assert 0 < rc["scores"].numel() <= cfg_c.model.post_process.nms.nms_post_max_size
    assert abs(rg["scores"].numel() - rc["scores"].numel()) <= max(2, rc["scores"].numel() // 50)
    k = min(rg["scores"].numel(), rc["scores"].numel(), 20)
    assert torch.allclose(rg["scores"][:k], rc["scores"][:k], atol=2e-3)
