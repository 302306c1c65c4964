"""Voxel-DETR (VD/voxel_detr.py:15-206): mean-VFE reader -> SparseResNet18 + FPN (p3) -> 1x1 conv +
GroupNorm -> box-attention encoder -> top-k proposals -> decoder -> per-layer detection heads;
training returns the loss dict, evaluation the decoded top-300 boxes per scene.

``model(batched_inputs)`` takes the reference's list of ``(point_voxels, info)`` pairs.  Two input
forms are accepted for ``point_voxels``:
  * the reference's CPU-voxelized dict (``voxels``, ``coordinates``, ``num_points_per_voxel``, ``shape``) —
    collated and copied to the device like efg/data/datasets/waymo/waymo.py:143-183;
  * a dict with only ``points`` (numpy [N,F] or a device tensor): the scene is voxelized on the GPU
    by the fused hash voxelizer + mean-VFE (bit-identical voxels, no CPU pass).
"""
import copy

import numpy as np
import torch
from torch import nn

from ... import ops
from ...backend import cuda_backend
from ...modeling.fpn import build_resnet_fpn_backbone
from ...modeling.voxel_reader import VoxelMeanFeatureExtractor
from .box_coder import VoxelBoxCoder3D
from .heads import Det3DHead
from .position_encoding import build_position_encoding
from .transformer import Transformer


class Backbone3d(nn.Module):
    def __init__(self, hidden_dim, reader, extractor, position_encoding, out_features=()):
        super().__init__()
        self.reader = reader
        self.extractor = extractor
        self.position_encoding = build_position_encoding(position_encoding, hidden_dim)
        self.out_features = list(out_features)
        self.num_channels = [extractor.out_channels] * len(self.out_features)

    def forward(self, voxels, coordinates, num_points_per_voxel, batch_size, input_shape):
        encoded = self.reader(voxels, num_points_per_voxel, coordinates)
        feats = self.extractor(encoded, coordinates, batch_size, input_shape)
        return [(feats[f], self.position_encoding(feats[f]).type_as(feats[f])) for f in self.out_features]


class _StaticSection(nn.Module):
    """Everything between the sparse backbone and the losses: FPN top-down convs -> input projection -> position
    encoding -> box-attention encoder -> top-k proposals -> decoder -> detection heads.  All of its shapes are fixed by
    the BEV grid, the batch size and num_queries, so forward AND backward can be held in CUDA graphs
    (``VoxelDETR.enable_static_graph``): ~1000 kernel launches of a launch-bound step become two graph launches.
    Not registered as a sub-module of the detector (its members already are): the state_dict is unchanged."""

    def __init__(self, det, feat_names):
        super().__init__()
        ext = det.backbone.extractor
        self.det = [det]
        self.feat_names = list(feat_names)
        self.lateral_convs = nn.ModuleList(ext.lateral_convs)
        self.output_convs = nn.ModuleList(ext.output_convs)
        self.input_proj = det.input_proj
        self.transformer = det.transformer

    def forward(self, *maps):
        det = self.det[0]
        feats = det.backbone.extractor.forward_dense(dict(zip(self.feat_names, maps)))
        feats_pos = [(feats[f], det.backbone.position_encoding(feats[f]).type_as(feats[f])) for f in det.backbone.out_features]
        features = [det.input_proj[i](fp[0]) for i, fp in enumerate(feats_pos)]
        hs, init_ref, inter_refs, memory, anchors, topk_idx = det.transformer(features, [fp[1] for fp in feats_pos])
        head = det.transformer.decoder.detection_head
        cls_out, box_out = [], []
        for i in range(hs.shape[0]):
            c, b = head(hs[i], init_ref if i == 0 else inter_refs[i - 1], i)
            cls_out.append(c)
            box_out.append(b)
        enc_cls, enc_box = det.transformer._enc_head_out
        det.transformer._enc_head_out = None
        return torch.stack(cls_out), torch.stack(box_out), memory, anchors, topk_idx, enc_cls, enc_box


def collate_voxels(samples, device):
    """The voxel part of waymo.py:143-183: concatenate per-scene arrays, prepend the batch index."""
    voxels = torch.from_numpy(np.concatenate([s["voxels"] for s in samples], 0)).to(device)
    npv = torch.from_numpy(np.concatenate([s["num_points_per_voxel"] for s in samples], 0)).to(device)
    coords = np.concatenate([np.pad(s["coordinates"], ((0, 0), (1, 0)), mode="constant", constant_values=i)
                             for i, s in enumerate(samples)], 0)
    return voxels, torch.from_numpy(coords).to(device), npv, np.asarray(samples[0]["shape"])


class VoxelDETR(nn.Module):
    def __init__(self, config, backend=None, prune_unused=True):
        super().__init__()
        self.backend = [backend or cuda_backend()]
        self.device = torch.device(config.model.device)
        self.hidden_dim = config.model.hidden_dim
        self.aux_loss = config.model.aux_loss
        self.num_classes = len(config.dataset.classes)
        self.num_queries = config.model.transformer.num_queries
        self.config = config
        self.stacked_losses = True  # decoder-layer losses / matching costs evaluated on layer-stacked tensors
        self.device_matching = True  # Hungarian assignments on the GPU (no host round trip); False = host scipy
        self.reuse_proposal_head = True  # False: evaluate the proposal head a second time for the loss, as the reference

        input_dim = len(config.dataset.format) if config.dataset.nsweeps == 1 else len(config.dataset.format) + 1
        self.input_dim = input_dim
        reader = VoxelMeanFeatureExtractor(**config.model.backbone.reader, num_input_features=input_dim)
        extractor = build_resnet_fpn_backbone(config.model.backbone.extractor, input_dim, backend=self.backend[0])
        if prune_unused:
            extractor.set_needed(config.model.backbone.out_features)
        self.backbone = Backbone3d(config.model.backbone.hidden_dim, reader, extractor,
                                   config.model.backbone.position_encoding,
                                   out_features=config.model.backbone.out_features)
        self.input_proj = nn.ModuleList([
            nn.Sequential(nn.Conv2d(c, self.hidden_dim, kernel_size=1), nn.GroupNorm(32, self.hidden_dim))
            for c in self.backbone.num_channels])
        for m in self.input_proj.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight, gain=1)
                nn.init.constant_(m.bias, 0)

        t = config.model.transformer
        self.transformer = self._build_transformer(config, t)
        self.transformer.proposal_head = Det3DHead(config, with_aux=False, with_metrics=False, num_classes=1,
                                                   num_layers=1)
        self.transformer.decoder.detection_head = Det3DHead(config, with_aux=True, with_metrics=True,
                                                            num_classes=self.num_classes, num_layers=t.dec_layers)
        self.box_coder = VoxelBoxCoder3D(config.dataset.voxel_size, config.dataset.pc_range, device=self.device)
        self._cpu_box_coder = VoxelBoxCoder3D(config.dataset.voxel_size, config.dataset.pc_range)
        grid = np.round((np.asarray(config.dataset.pc_range[3:], dtype=np.float32) -
                         np.asarray(config.dataset.pc_range[:3], dtype=np.float32)) /
                        np.asarray(config.dataset.voxel_size, dtype=np.float32)).astype(np.int64)
        self.grid_size = grid  # (x, y, z)
        self.to(self.device)

    def _build_transformer(self, config, t):
        return Transformer(d_model=t.hidden_dim, nhead=t.nhead, nlevel=len(config.model.backbone.out_features),
                           num_encoder_layers=t.enc_layers, num_decoder_layers=t.dec_layers,
                           dim_feedforward=t.dim_feedforward, dropout=t.dropout, num_queries=t.num_queries,
                           backend=self.backend[0])

    # ---------------------------------------------------------------------------------------
    def voxelize_on_device(self, samples):
        """GPU path: samples hold raw ``points``; one fused voxelize + mean-VFE launch sequence for the batch."""
        ds = self.config.dataset
        pts, sizes = [], []
        for s in samples:
            p = s["points"]
            if not isinstance(p, torch.Tensor):
                p = torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32))
            pts.append(p.to(self.device, non_blocking=True))
            sizes.append(p.shape[0])
        points = torch.cat(pts, 0) if len(pts) > 1 else pts[0]
        offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
        if self.device.type == "cuda":
            offs = offs.pin_memory()
        offs = offs.to(self.device, non_blocking=True)
        max_voxels = ds.get("max_voxel_num", 150000)
        if isinstance(max_voxels, (list, tuple)):
            max_voxels = max_voxels[0] if self.training else max_voxels[1]
        r = ops.hard_voxelize_batched(points.contiguous(), offs, ds.voxel_size, ds.pc_range,
                                      ds.get("max_points_in_voxel", 5), max_voxels, coors_dim=4, want_voxels=False,
                                      want_mean=True)
        m = int(r["counts"][-1].item())  # one host sync: the voxel count sizes every later tensor
        return r["mean"][:m], r["coors"][:m], r["num_points_per_voxel"][:m], self.grid_size

    def encode_targets(self, batched_inputs):
        """Ground truth of the batch -> normalised box codes (box_coder.encode) and 0-based labels.  The
        encoding runs on the host copy (a few hundred boxes) and the batch is uploaded as TWO pinned,
        non-blocking copies; per-scene dicts are views into the concatenation."""
        from .losses import TargetList

        cpu_coder = self._cpu_box_coder
        enc = []
        for _, info in batched_inputs:
            ann = info["annotations"]
            t = {"gt_boxes": torch.as_tensor(np.asarray(ann["gt_boxes"], dtype=np.float32)),
                 "labels": torch.as_tensor(np.asarray(ann["labels"], dtype=np.int64))}
            enc.append(cpu_coder.encode(t))
        sizes = [int(e["labels"].shape[0]) for e in enc]
        boxes = torch.cat([e["gt_boxes"] for e in enc], 0) if enc else torch.zeros((0, 7))
        labels = torch.cat([e["labels"] for e in enc], 0) if enc else torch.zeros((0,), dtype=torch.int64)
        if self.device.type == "cuda":
            boxes = boxes.pin_memory().to(self.device, non_blocking=True)
            labels = labels.pin_memory().to(self.device, non_blocking=True)
        targets = TargetList()
        offsets, off = [], 0
        for n in sizes:
            targets.append({"gt_boxes": boxes[off:off + n], "labels": labels[off:off + n]})
            offsets.append(off)
            off += n
        targets.labels_cat, targets.boxes_cat, targets.offsets = labels, boxes, offsets
        return targets

    def extract(self, batched_inputs, prepared=None):
        batch_size = len(batched_inputs)
        if prepared is not None:
            feats = self.backbone.extractor.forward_dense(self.bottom_up_maps(batched_inputs, prepared))
            feats_pos = [(feats[f], self.backbone.position_encoding(feats[f]).type_as(feats[f])) for f in self.backbone.out_features]
            return [self.input_proj[i](fp[0]) for i, fp in enumerate(feats_pos)], [fp[1] for fp in feats_pos]
        samples = [bi[0] for bi in batched_inputs]
        if "voxels" in samples[0]:
            voxels, coords, npv, input_shape = collate_voxels(samples, self.device)
        else:
            voxels, coords, npv, input_shape = self.voxelize_on_device(samples)
        feats_pos = self.backbone(voxels, coords, npv, batch_size, input_shape)
        features = [self.input_proj[i](fp[0]) for i, fp in enumerate(feats_pos)]
        return features, [fp[1] for fp in feats_pos]

    # ---------------------------------------------------------------------------------------
    def prepare(self, batched_inputs, stream=None):
        """The index part of a step for raw-point samples, without features or parameters: host-to-device copy of the
        points, voxelizer (+ mean VFE) and every strided rulebook of the sparse backbone.  These are the only places of
        a training step that read a count back from the device; a loader calls this for batch i + 1 on its own (high
        priority) stream while step i runs, and passes the result to forward(..., prepared=...), which then never
        synchronises: the host runs ahead of the GPU instead of draining it once per downsampling."""
        samples = [bi[0] for bi in batched_inputs]
        if "voxels" in samples[0] or self.device.type != "cuda":
            return None
        bottom_up = self.backbone.extractor.bottom_up
        if not hasattr(bottom_up, "plan_geometry"):
            return None
        stream = stream or torch.cuda.current_stream(self.device)
        with torch.cuda.stream(stream):
            voxels, coords, npv, input_shape = self.voxelize_on_device(samples)
            indice_dict = bottom_up.plan_geometry(coords, len(batched_inputs), input_shape)
            done = torch.cuda.Event()
            done.record(stream)
        return {"voxels": voxels, "coords": coords, "npv": npv, "input_shape": input_shape, "indice_dict": indice_dict,
                "done": done, "stream": stream, "batch_size": len(batched_inputs)}

    @staticmethod
    def _adopt(prepared):
        """Make the tensors of prepare() safe to use on the current stream: order after the producer stream and tell
        the caching allocator about the second stream."""
        cur = torch.cuda.current_stream()
        if prepared["stream"] == cur:
            return
        cur.wait_event(prepared["done"])
        tensors = [prepared["voxels"], prepared["coords"], prepared["npv"]]
        for rb in prepared["indice_dict"].values():
            tensors += [t for t in (rb.nbr, rb.nbr_t, rb.out_indices) if isinstance(t, torch.Tensor)]
        for t in tensors:
            t.record_stream(cur)
        prepared["stream"] = cur

    def bottom_up_maps(self, batched_inputs, prepared=None):
        """Voxelize + sparse backbone: the dense bottom-up feature maps the FPN needs (dynamic shapes end here)."""
        batch_size = len(batched_inputs)
        samples = [bi[0] for bi in batched_inputs]
        if prepared is not None:
            assert prepared["batch_size"] == batch_size
            self._adopt(prepared)
            voxels, coords, npv, input_shape = (prepared[k] for k in ("voxels", "coords", "npv", "input_shape"))
            encoded = self.backbone.reader(voxels, npv, coords)
            return self.backbone.extractor.bottom_up(encoded, coords, batch_size, input_shape, indice_dict=prepared["indice_dict"])
        if "voxels" in samples[0]:
            voxels, coords, npv, input_shape = collate_voxels(samples, self.device)
        else:
            voxels, coords, npv, input_shape = self.voxelize_on_device(samples)
        encoded = self.backbone.reader(voxels, npv, coords)
        return self.backbone.extractor.bottom_up(encoded, coords, batch_size, input_shape)

    def enable_static_graph(self, batched_inputs):
        """Capture the static section (see _StaticSection) into CUDA graphs, forward and backward
        (torch.cuda.make_graphed_callables), using `batched_inputs` as the sample.  Training mode, fixed batch size.
        Returns True on success; on failure the model keeps running eagerly and the reason is kept in
        ``self.static_graph_error`` (never silent: bench.py reports it)."""
        self.static_graph_error = None
        try:
            if not (self.training and self.device.type == "cuda" and self.reuse_proposal_head):
                raise RuntimeError("needs a CUDA model in training mode with reuse_proposal_head")
            feats = self.bottom_up_maps(batched_inputs)
            names = [n for n in self.backbone.extractor.in_features if n in feats]
            section = _StaticSection(self, names)
            sample = tuple(feats[n].detach().clone().requires_grad_(True) for n in names)
            torch.cuda.synchronize()
            # weight images: the captured kernels read the cached images, which forward() refreshes in one launch
            # before every replay (ops.refresh_packs) instead of one captured pack kernel per layer and direction
            ops.PACKS_REFRESHED_PER_STEP = True
            count0 = ops.launch_count()
            try:
                self._static_call = torch.cuda.make_graphed_callables(section, sample, allow_unused_input=True)
            finally:
                ops.PACKS_REFRESHED_PER_STEP = False
            # library kernels one replay (forward + backward graph) executes: make_graphed_callables runs the section
            # three times eagerly and once under capture
            self.static_graph_launches = (ops.launch_count() - count0) // 4
            self._static_packs = ops.pin_pack_cache()   # the graphs hold raw pointers into these images
            self._static_names, self._static_batch = names, len(batched_inputs)
            self._static_section = [section]
            return True
        except Exception as e:  # noqa: BLE001 — capture failures are reported, the eager path stays intact
            self._static_call = None
            self.static_graph_error = "%s: %s" % (type(e).__name__, e)
            return False

    def forward(self, batched_inputs, prepared=None):
        """`prepared`: the result of prepare(batched_inputs) (optional; see there)."""
        if self.training and self.device.type == "cuda":
            ops.refresh_packs()   # every weight image the optimizer step made stale, in one launch
        targets = self.encode_targets(batched_inputs) if self.training else None
        if targets is not None:
            # the loss normaliser depends on the targets only: its all-reduce runs under the forward pass
            self.transformer.proposal_head.losses.request_normaliser(targets, self.device)
        call = getattr(self, "_static_call", None)
        if call is not None and self.training and torch.is_grad_enabled() and len(batched_inputs) == self._static_batch:
            feats = self.bottom_up_maps(batched_inputs, prepared)
            cls_out, box_out, memory, anchors, topk_idx, enc_cls, enc_box = call(*[feats[n] for n in self._static_names])
            self.transformer._enc_head_out = (enc_cls, enc_box)
            return self.losses(cls_out, box_out, memory, anchors, topk_idx, targets)
        features, pos = self.extract(batched_inputs, prepared)
        hs, init_ref, inter_refs, memory, anchors, topk_idx = self.transformer(features, pos)

        head = self.transformer.decoder.detection_head
        cls_out, box_out = [], []
        for i in range(hs.shape[0]):
            ref = init_ref if i == 0 else inter_refs[i - 1]
            c, b = head(hs[i], ref, i)
            cls_out.append(c)
            box_out.append(b)
        cls_out, box_out = torch.stack(cls_out), torch.stack(box_out)

        if self.training:
            return self.losses(cls_out, box_out, memory, anchors, topk_idx, targets)
        return self.postprocess(cls_out[-1], box_out[-1])

    def losses(self, cls_out, box_out, memory, anchors, topk_idx, targets):
        """Encoder-proposal and decoder losses.  The cost matrices of all four matchings (encoder + 3
        decoder layers) are built on the GPU and solved after ONE device-to-host transfer; the assignments
        go back in one pinned non-blocking copy (the reference syncs once per layer and per index tensor,
        VD/modules/matcher.py:86, VD/losses.py:17-48)."""
        from .losses import TargetList, device_matches, upload_matches

        losses = {}
        prop, head = self.transformer.proposal_head, self.transformer.decoder.detection_head
        num_boxes = prop.losses.normaliser(targets, cls_out.device)
        # the reference evaluates the proposal head twice on the same (memory, anchors) — once, detached, to pick the
        # top-k proposals (VD/transformer.py:56) and once for this loss (VD/voxel_detr.py:145); the transformer keeps
        # the first result, which is the same tensor with its graph attached
        cached = getattr(self.transformer, "_enc_head_out", None) if self.reuse_proposal_head else None
        enc_cls, enc_box = cached if cached is not None else prop(memory, anchors)
        self.transformer._enc_head_out = None
        bin_targets = TargetList(dict(t, labels=torch.zeros_like(t["labels"])) for t in targets)
        bin_targets.labels_cat = torch.zeros_like(targets.labels_cat)
        bin_targets.boxes_cat, bin_targets.offsets = targets.boxes_cat, targets.offsets
        enc_out = {"topk_indexes": topk_idx, "pred_logits": enc_cls, "pred_boxes": enc_box}
        dec_out = {"pred_logits": cls_out[-1], "pred_boxes": box_out[-1],
                   "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(cls_out[:-1], box_out[:-1])]}
        stacked = self.stacked_losses and isinstance(targets, TargetList) and targets.labels_cat is not None
        if stacked:
            dec_mats = head.losses.matcher.cost_matrices_stacked(cls_out, box_out, targets)
        else:
            dec_mats = head.losses.prepare(dec_out, targets)
        mats = prop.losses.prepare(enc_out, bin_targets) + dec_mats
        matcher = head.losses.matcher
        if self.device_matching and mats and mats[0].is_cuda and "solve" not in matcher.__dict__:
            # assignments solved on the device (scipy's algorithm, csrc/lsa.cu): the step has no host round trip here
            matches = device_matches(mats, len(mats) // max(len(targets), 1), targets.offsets, cls_out.device)
        else:
            solved = matcher.solve(mats)  # host scipy, one device-to-host transfer (the reference's path)
            bs = len(targets)
            per_layer = [solved[i * bs:(i + 1) * bs] for i in range(len(solved) // max(bs, 1))]
            matches = upload_matches(per_layer, targets.offsets, cls_out.device)
        enc = prop.compute_losses(enc_out, bin_targets, num_boxes, solved=matches[:1])
        losses.update({k + "_enc": v for k, v in enc.items()})
        losses.update(head.compute_losses(dec_out, targets, num_boxes, solved=matches[1:],
                                          stacked=(cls_out, box_out) if stacked else None))
        return losses

    def postprocess(self, logits, boxes):
        """sigmoid -> top-300 over (query, class) -> decode (VD/voxel_detr.py:168-198)."""
        prob = logits.sigmoid().view(logits.shape[0], -1)
        boxes = self.box_coder.decode(boxes)
        k = min(300, prob.shape[1])
        scores, idx = torch.topk(prob, k, dim=1, sorted=False)
        q = idx.div(logits.shape[2], rounding_mode="floor")
        labels = idx % logits.shape[2] + 1
        picked = torch.gather(boxes, 1, q.unsqueeze(-1).repeat(1, 1, boxes.shape[-1]))
        return [{"scores": s.detach().cpu(), "labels": l.detach().cpu(), "boxes3d": b.detach().cpu()}
                for s, l, b in zip(scores, labels, picked)]


def build_model(self, config, backend=None):
    """Plugin entry point, same signature as the playground's ``net.build_model(self, config)``."""
    return VoxelDETR(config, backend=backend)
