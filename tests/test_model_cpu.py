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


def test_conquer_cpu_oracle_train_step():
    from efg_b200.config import conquer_config
    from efg_b200.detectors.conquer import ConQueR

    torch.manual_seed(0)
    cfg = conquer_config(
        dataset={"pc_range": SMALL.pc_range, "voxel_size": SMALL.voxel_size, "max_voxel_num": 20000},
        model={"device": "cpu", "transformer": {"num_queries": 40, "enc_layers": 1, "dec_layers": 2}})
    model = ConQueR(cfg, backend=cpu_backend()).train()
    before = [p.detach().clone() for p in model.transformer.decoder_gt.parameters()]
    batch = [(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in small_batch()]
    losses = model(batch)
    for k in ("loss_ce", "loss_ce_dn", "loss_bbox_dn_0", "loss_giou_dn", "loss_contrastive_dec_0",
              "loss_contrastive_dec_1", "loss_ce_enc"):
        assert k in losses, (k, sorted(losses))
    total = sum(v for k, v in losses.items() if k.startswith("loss"))
    assert torch.isfinite(total)
    total.backward()
    assert all(p.grad is None for p in model.transformer.decoder_gt.parameters())  # EMA copy takes no gradient
    assert model.projector[0].weight.grad is not None and model.predictor[2].weight.grad is not None
    # the EMA copy starts equal to the decoder, so one momentum step leaves it unchanged up to rounding
    for b, a in zip(before, model.transformer.decoder_gt.parameters()):
        assert torch.allclose(a, b, atol=1e-6)
    model.eval()
    out = model(batch[:1])
    assert set(out[0]) == {"scores", "labels", "boxes3d"}


def test_conquer_contrastive_matches_reference_loops():
    """The vectorised InfoNCE equals the reference's per-pair Python loops (CQ/voxel_detr.py:227-254)."""
    torch.manual_seed(3)
    B, nq, groups, max_gt, tau = 2, 12, 3, 4, 0.7
    per_gt = [4, 2]
    gt_projs = torch.randn(B, (groups + 1) * max_gt, 8)
    pred_projs = torch.randn(B, nq, 8)
    matched = [(torch.tensor([3, 7, 1, 9]), torch.tensor([0, 1, 2, 3])), (torch.tensor([5, 0]), torch.tensor([1, 0]))]
    sim_f = torch.nn.CosineSimilarity(dim=2)
    ref = 0.0
    for bi, (src, tgt) in enumerate(matched):
        sim = sim_f(gt_projs[bi].unsqueeze(1), pred_projs[bi].unsqueeze(0)) / tau
        neg_mask = torch.ones(nq, dtype=torch.bool)
        neg_mask[src] = False
        for q, t in zip(src.tolist(), tgt.tolist()):
            pos_mask = torch.tensor([t + max_gt * pi for pi in range(1, groups + 1)])
            pos_pair = sim[pos_mask, q].view(-1, 1)
            neg_pairs = sim[:, neg_mask][pos_mask]
            ref = ref + (torch.log(torch.exp(pos_pair) + torch.exp(neg_pairs).sum(-1, keepdim=True)) - pos_pair).mean()
    # vectorised form used by ConQueR.contrastive_losses
    b_idx = torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(matched)])
    q_idx = torch.cat([s for s, _ in matched])
    local_t = torch.cat([t for _, t in matched])
    neg = torch.ones(B, nq, dtype=torch.bool)
    neg[b_idx, q_idx] = False
    pos_rows = local_t[:, None] + max_gt * torch.arange(1, groups + 1)[None, :]
    g = torch.nn.functional.normalize(gt_projs, dim=-1, eps=1e-8)
    p = torch.nn.functional.normalize(pred_projs, dim=-1, eps=1e-8)
    sim = torch.einsum("bgc,bqc->bgq", g, p) / tau
    rows = sim[b_idx[:, None], pos_rows]
    pos = torch.gather(rows, 2, q_idx[:, None, None].expand(-1, groups, 1))
    negsum = (torch.exp(rows) * neg[b_idx][:, None, :]).sum(-1, keepdim=True)
    vec = (torch.log(torch.exp(pos) + negsum) - pos).mean(dim=1).sum()
    assert torch.allclose(vec, ref, atol=1e-5)


def test_centerpoint_cpu_oracle_train_step():
    from efg_b200.config import centerpoint_config
    from efg_b200.detectors.centerpoint import VoxelNet

    torch.manual_seed(0)
    cfg = centerpoint_config(dataset={"pc_range": SMALL.pc_range, "voxel_size": SMALL.voxel_size, "max_voxel_num": 20000},
                             model={"device": "cpu"})
    model = VoxelNet(cfg, backend=cpu_backend()).train()
    batch = [(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in small_batch(2, 5000, seed=2)]
    losses = model(batch)
    assert {"0_loss", "0_hm_loss", "0_loc_loss", "0_num_positive"} == set(losses)
    assert float(losses["0_num_positive"]) > 0
    losses["0_loss"].backward()
    g = dict(model.named_parameters())["backbone.conv_input.0.weight"].grad
    assert g is not None and torch.isfinite(g).all()
    # legacy SparseBasicBlock convs carry a bias (sparse_net.py:443-448), the strided convs do not
    sd = model.state_dict()
    assert "backbone.conv1.0.conv1.bias" in sd and "backbone.conv2.0.bias" not in sd
    assert sd["backbone.conv4.0.weight"].shape == (128, 3, 3, 3, 64)
    model.eval()
    with torch.no_grad():
        out = model(batch[:1])
    assert set(out[0]) == {"boxes3d", "scores", "labels"} and out[0]["boxes3d"].shape[1] == 7


def test_centerpoint_label_assignment_heatmap():
    from efg_b200.detectors.centerpoint.assign import assign_scene

    ann = {"gt_boxes": np.array([[0.0, 0.0, 0.0, 4.0, 2.0, 1.5, 0, 0, 0.3], [5.0, -3.0, 0.0, 0.8, 0.8, 1.7, 0, 0, -1.0]], np.float32),
           "gt_names": np.array(["VEHICLE", "PEDESTRIAN"])}
    tasks = [{"num_classes": 3, "class_names": ["VEHICLE", "PEDESTRIAN", "CYCLIST"]}]
    t = assign_scene(ann, tasks, np.array([256, 256, 40]), SMALL.pc_range, SMALL.voxel_size, 8, 0.1, 500, 2)
    hm = t["hm"][0]
    assert hm.shape == (3, 32, 32) and hm[0].max() == 1.0 and hm[1].max() == 1.0 and hm[2].max() == 0.0
    assert t["mask"][0].sum() == 2
    y, x = divmod(int(t["ind"][0][0]), 32)
    assert hm[0, y, x] == 1.0  # the peak sits on the object's centre cell
    assert np.allclose(t["anno_box"][0][0, 3:6], np.log([4.0, 2.0, 1.5]), atol=1e-6)


def test_stacked_losses_equal_per_layer_reference_path():
    """Evaluating matching costs and losses of all decoder layers on layer-stacked tensors (and reusing the
    proposal head's output for the encoder loss) is the reference's layer-by-layer evaluation
    (VD/losses.py:98-141, VD/voxel_detr.py:141-166) with fewer kernel launches: same losses, same gradients."""
    torch.manual_seed(3)
    cfg = small_config()
    ref = VoxelDETR(cfg, backend=cpu_backend())
    fast = VoxelDETR(cfg, backend=cpu_backend())
    fast.load_state_dict(ref.state_dict())
    ref.stacked_losses = False
    ref.reuse_proposal_head = False  # the reference evaluates the proposal head a second time
    ref.train()
    fast.train()
    batch = [(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in small_batch(2, 4000, seed=11)]
    out = []
    for m in (ref, fast):
        losses = m(batch)
        total = sum(v for k, v in losses.items() if k.startswith("loss"))
        total.backward()
        out.append((losses, {n: p.grad for n, p in m.named_parameters() if p.grad is not None}))
    (l0, g0), (l1, g1) = out
    assert set(l0) == set(l1)
    for k in l0:
        assert abs(float(l0[k]) - float(l1[k])) <= 1e-5 * max(1.0, abs(float(l0[k]))), (k, float(l0[k]), float(l1[k]))
    assert set(g0) == set(g1)
    for n in g0:
        scale = max(1.0, g0[n].abs().max().item())
        assert (g0[n] - g1[n]).abs().max().item() <= 2e-4 * scale, n


def test_losses_with_scenes_without_ground_truth():
    """One scene of the batch (and then the whole batch) has no boxes: both loss evaluations must agree and stay
    finite (the classification loss still sees every query as background)."""
    torch.manual_seed(4)
    cfg = small_config()
    model = VoxelDETR(cfg, backend=cpu_backend()).train()
    scenes = small_batch(2, 4000, seed=17)
    empty = {k: v[:0] for k, v in scenes[1][1].items()}
    for anns in ([scenes[0][1], empty], [empty, empty]):
        batch = [(voxelized_sample(p, cfg.dataset), {"annotations": a}) for (p, _), a in zip(scenes, anns)]
        res = []
        for stacked in (False, True):
            model.stacked_losses = stacked
            losses = model(batch)
            total = sum(v for k, v in losses.items() if k.startswith("loss"))
            assert torch.isfinite(total)
            res.append({k: float(v) for k, v in losses.items()})
        for k in res[0]:
            assert abs(res[0][k] - res[1][k]) <= 1e-5 * max(1.0, abs(res[0][k])), k
