"""CPU: the Voxel-DETR module graph wired over the CPU oracle backend trains one step on a tiny
scene (host logic: collate, targets, matcher, losses, state_dict layout)."""
import numpy as np
import torch

from efg_b200.config import voxel_detr_config
from efg_b200.data import SceneSpec, make_scene
from efg_b200.detectors.voxel_detr import VoxelDETR
from oracle.backend_cpu import cpu_backend, voxelized_sample

SMALL = SceneSpec(pc_range=[-12.8, -12.8, -2.0, 12.8, 12.8, 4.0], voxel_size=[0.1, 0.1, 0.15])


def small_config(device="cpu", num_queries=40):
    return voxel_detr_config(
        dataset={"pc_range": SMALL.pc_range, "voxel_size": SMALL.voxel_size, "max_voxel_num": 20000},
        model={"device": device, "transformer": {"num_queries": num_queries, "enc_layers": 1, "dec_layers": 2}})


def small_batch(n_scenes=2, n_points=4000, seed=0):
    out = []
    for i in range(n_scenes):
        pts, ann = make_scene(n_points, SMALL, seed=seed + i, num_objects=6)
        keep = (np.abs(ann["gt_boxes"][:, 0]) < 12) & (np.abs(ann["gt_boxes"][:, 1]) < 12)
        ann = {k: v[keep] for k, v in ann.items()}
        out.append((pts, ann))
    return out


def test_voxel_detr_cpu_oracle_train_step():
    torch.manual_seed(0)
    cfg = small_config()
    model = VoxelDETR(cfg, backend=cpu_backend())
    model.train()
    batch = [(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in small_batch()]
    losses = model(batch)
    expected = {"loss_ce", "loss_bbox", "loss_giou", "loss_rad", "loss_ce_0", "loss_bbox_0", "loss_giou_0",
                "loss_rad_0", "loss_ce_enc", "loss_bbox_enc", "loss_giou_enc", "loss_rad_enc", "accuracy"}
    assert expected == set(losses.keys())
    total = sum(v for k, v in losses.items() if k.startswith("loss"))
    assert torch.isfinite(total)
    total.backward()
    grads = {n: p.grad for n, p in model.named_parameters()}
    assert grads["backbone.extractor.bottom_up.stem.conv1.0.weight"] is not None
    assert torch.isfinite(grads["backbone.extractor.bottom_up.stem.conv1.0.weight"]).all()
    # pruned FPN branches keep their parameters but receive no gradient (find_unused_parameters in the reference)
    assert grads["backbone.extractor.fpn_output2.weight"] is None
    assert grads["backbone.extractor.fpn_output3.weight"] is not None


def test_voxel_detr_pruned_equals_full_graph():
    """Skipping the FPN levels nobody reads is output-identical to the reference's full evaluation."""
    torch.manual_seed(1)
    cfg = small_config()
    full = VoxelDETR(cfg, backend=cpu_backend(), prune_unused=False)
    pruned = VoxelDETR(cfg, backend=cpu_backend(), prune_unused=True)
    pruned.load_state_dict(full.state_dict())
    batch = [(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in small_batch(1, 3000, seed=5)]
    full.eval()
    pruned.eval()
    with torch.no_grad():
        fa, _ = full.extract(batch)
        fb, _ = pruned.extract(batch)
    assert torch.equal(fa[0], fb[0])
    res = pruned(batch)
    assert res[0]["boxes3d"].shape[1] == 7 and res[0]["labels"].min() >= 1


def test_box_coder_roundtrip():
    from efg_b200.detectors.voxel_detr.box_coder import VoxelBoxCoder3D

    coder = VoxelBoxCoder3D(SMALL.voxel_size, SMALL.pc_range)
    boxes = torch.tensor([[1.0, -2.0, 0.5, 4.0, 2.0, 1.5, 0.0, 0.0, 2.5], [-5.0, 7.0, -1.0, 1.0, 1.0, 1.7, 0, 0, -3.0]])
    t = coder.encode({"gt_boxes": boxes.clone(), "labels": torch.tensor([1, 3])})
    assert t["labels"].tolist() == [0, 2] and t["gt_boxes"].shape == (2, 7)
    back = coder.decode(t["gt_boxes"])
    assert torch.allclose(back[:, :6], boxes[:, :6], atol=1e-5)
    assert torch.allclose(torch.cos(back[:, 6]), torch.cos(boxes[:, 8]), atol=1e-5)
