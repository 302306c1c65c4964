"""CenterPoint ``VoxelNet`` (CP/voxelnet.py:19-226): mean-VFE reader -> SpMiddleResNetFHD sparse encoder ->
RPN neck -> CenterHead; training returns the per-task loss dict.  Inputs as for Voxel-DETR: the
reference's CPU-voxelized sample dicts, or raw ``points`` that are voxelized on the GPU."""
import numpy as np
import torch
from torch import nn

from ... import ops
from ...backend import cuda_backend
from ...modeling.rpn import RPN
from ...modeling.sparse_backbone import SpMiddleResNetFHD
from ...modeling.voxel_reader import VoxelMeanFeatureExtractor
from ..voxel_detr.model import collate_voxels
from .assign import assign_scene
from .head import CenterHead


class VoxelNet(nn.Module):
    def __init__(self, config, backend=None):
        super().__init__()
        self.config = config
        self.backend = [backend or cuda_backend()]
        self.device = torch.device(config.model.device)
        self.reader = VoxelMeanFeatureExtractor(**config.model.reader)
        self.backbone = SpMiddleResNetFHD(**config.model.backbone, backend=self.backend[0])
        self.neck = RPN(config.model.neck)
        self.center_head = CenterHead(config)
        self.center_head.rotate_nms = getattr(self.backend[0], "rotate_nms", None)
        a = config.model.loss
        self.out_size_factor = a.out_size_factor
        self.tasks = config.model.head.tasks
        self.gaussian_overlap, self._max_objs, self._min_radius = a.gaussian_overlap, a.max_objs, a.min_radius
        pr = np.asarray(config.dataset.pc_range, dtype=np.float32)
        vs = np.asarray(config.dataset.voxel_size, dtype=np.float32)
        self.grid_size = np.round((pr[3:] - pr[:3]) / vs).astype(np.int64)
        self.to(self.device)

    def voxelize_on_device(self, samples):
        ds = self.config.dataset
        pts = [s["points"] if isinstance(s["points"], torch.Tensor) else
               torch.from_numpy(np.ascontiguousarray(s["points"], dtype=np.float32)) for s in samples]
        sizes = [p.shape[0] for p in pts]
        points = torch.cat([p.to(self.device, non_blocking=True) for p in pts], 0)
        offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
        if self.device.type == "cuda":
            offs = offs.pin_memory()
        offs = offs.to(self.device, non_blocking=True)
        r = ops.hard_voxelize_batched(points.contiguous(), offs, ds.voxel_size, ds.pc_range,
                                      ds.get("max_points_in_voxel", 5), ds.get("max_voxel_num", 150000), coors_dim=4,
                                      want_voxels=False, want_mean=True)
        m = int(r["counts"][-1].item())
        return r["mean"][:m], r["coors"][:m], r["num_points_per_voxel"][:m], self.grid_size

    def label_assign(self, infos):
        ds = self.config.dataset
        if self.device.type == "cuda" and getattr(self.backend[0], "name", "") == "efgb200-cuda":
            from .assign import assign_batch_device

            return assign_batch_device(infos, self.tasks, self.grid_size, ds.pc_range, ds.voxel_size, self.out_size_factor,
                                       self.gaussian_overlap, self._max_objs, self._min_radius, self.device)
        per_scene = [assign_scene(info["annotations"], self.tasks, self.grid_size, ds.pc_range, ds.voxel_size,
                                  self.out_size_factor, self.gaussian_overlap, self._max_objs, self._min_radius)
                     for info in infos]
        targets = {}
        for key in ("hm", "anno_box", "ind", "mask", "cat"):
            targets[key] = []
            for t in range(len(self.tasks)):
                arr = torch.from_numpy(np.stack([s[key][t] for s in per_scene], axis=0))
                if self.device.type == "cuda":
                    arr = arr.pin_memory()
                targets[key].append(arr.to(self.device, non_blocking=True))
        return targets

    # ---------------------------------------------------------------------------------------
    _TARGET_KEYS = ("hm", "anno_box", "ind", "mask", "cat")

    def prepare(self, batched_inputs, stream=None):
        """The index part of a training step for raw-point samples (see VoxelDETR.prepare): host-to-device copy of the
        points, voxelizer (+ mean VFE), the strided rulebooks of the sparse encoder and the label assignment — everything
        without parameters, including the only device-to-host reads of a step.  Run it for batch i + 1 on a side stream
        while step i executes and pass the result to forward(..., prepared=...)."""
        samples = [bi[0] for bi in batched_inputs]
        if "voxels" in samples[0] or self.device.type != "cuda" or not hasattr(self.backbone, "plan_geometry"):
            return None
        stream = stream or torch.cuda.current_stream(self.device)
        with torch.cuda.stream(stream), torch.no_grad():
            voxels, coords, npv, input_shape = self.voxelize_on_device(samples)
            indice_dict = self.backbone.plan_geometry(coords, len(batched_inputs), input_shape)
            if indice_dict is None:
                return None
            targets = self.label_assign([bi[1] for bi in batched_inputs]) if self.training else None
            done = torch.cuda.Event()
            done.record(stream)
        return {"voxels": voxels, "coords": coords, "npv": npv, "input_shape": input_shape, "indice_dict": indice_dict,
                "targets": targets, "done": done, "stream": stream, "batch_size": len(batched_inputs)}

    @staticmethod
    def _adopt(prepared):
        cur = torch.cuda.current_stream()
        if prepared["stream"] == cur:
            return
        cur.wait_event(prepared["done"])
        tensors = [prepared["voxels"], prepared["coords"], prepared["npv"]]
        for rb in prepared["indice_dict"].values():
            tensors += [t for t in (rb.nbr, rb.nbr_t, rb.out_indices) if isinstance(t, torch.Tensor)]
        if prepared["targets"] is not None:
            for vals in prepared["targets"].values():
                tensors += list(vals)
        for t in tensors:
            t.record_stream(cur)
        prepared["stream"] = cur

    def bev_map(self, batched_inputs, prepared=None):
        """Voxelizer + sparse encoder: the dense BEV map the neck consumes (dynamic shapes end here)."""
        samples = [bi[0] for bi in batched_inputs]
        batch_size = len(samples)
        if prepared is not None:
            assert prepared["batch_size"] == batch_size
            self._adopt(prepared)
            voxels, coords, npv, input_shape = (prepared[k] for k in ("voxels", "coords", "npv", "input_shape"))
            return self.backbone(self.reader(voxels, npv), coords, batch_size, input_shape, indice_dict=prepared["indice_dict"])
        with torch.no_grad():
            if "voxels" in samples[0]:
                voxels, coords, npv, input_shape = collate_voxels(samples, self.device)
            else:
                voxels, coords, npv, input_shape = self.voxelize_on_device(samples)
        return self.backbone(self.reader(voxels, npv), coords, batch_size, input_shape)

    def enable_static_graph(self, batched_inputs):
        """Capture neck + heads + losses, forward and backward, into CUDA graphs (torch.cuda.make_graphed_callables): every
        shape after the sparse encoder is fixed by the BEV grid, the batch size and max_objs.  Training mode, fixed batch
        size, the job on a non-default stream.  Returns True on success; on failure the model keeps running eagerly and
        the reason is kept in ``self.static_graph_error``."""
        self.static_graph_error = None
        try:
            if not (self.training and self.device.type == "cuda"):
                raise RuntimeError("needs a CUDA model in training mode")
            bev = self.bev_map(batched_inputs)
            with torch.no_grad():
                targets = self.label_assign([bi[1] for bi in batched_inputs])
            section = _StaticSection(self)
            flat = [t for k in self._TARGET_KEYS for t in targets[k]]
            sample = (bev.detach().clone().requires_grad_(True),) + tuple(t.clone() for t in flat)
            torch.cuda.synchronize()
            count0 = ops.launch_count()
            self._static_call = torch.cuda.make_graphed_callables(section, sample, allow_unused_input=True)
            self.static_graph_launches = (ops.launch_count() - count0) // 4   # library kernels per replay (none: cuDNN / ATen only)
            self._static_section = [section]
            self._static_batch = len(batched_inputs)
            return True
        except Exception as e:  # noqa: BLE001 — capture failures are reported, the eager path stays intact
            self._static_call = None
            self.static_graph_error = "%s: %s" % (type(e).__name__, e)
            return False

    def forward(self, batched_inputs, prepared=None):
        """`prepared`: the result of prepare(batched_inputs) (optional; see there)."""
        if self.training and self.device.type == "cuda":
            ops.refresh_packs()   # every weight image the optimizer step made stale, in one launch
        infos = [bi[1] for bi in batched_inputs]
        x = self.bev_map(batched_inputs, prepared)
        call = getattr(self, "_static_call", None)
        if self.training:
            if prepared is not None and prepared["targets"] is not None:
                targets = prepared["targets"]
            else:
                with torch.no_grad():
                    targets = self.label_assign(infos)
            if call is not None and torch.is_grad_enabled() and len(batched_inputs) == self._static_batch:
                section = self._static_section[0]
                values = call(x, *[t for k in self._TARGET_KEYS for t in targets[k]])
                return dict(zip(section.keys, values))
        preds = self.center_head(self.neck(x))
        if self.training:
            return self.center_head.loss(targets, preds)
        return self.center_head.decode(preds, self.config.model.post_process)


class _StaticSection(nn.Module):
    """Neck + heads + losses of CenterPoint as one callable over tensors (bev map, flattened targets) -> tuple of loss
    terms: the part of a training step whose shapes never change.  Not registered as a sub-module of the detector."""

    def __init__(self, det):
        super().__init__()
        self.det = [det]
        self.neck = det.neck
        self.center_head = det.center_head
        self.keys = None

    def forward(self, bev, *flat):
        det = self.det[0]
        tasks = len(det.tasks)
        targets = {k: list(flat[i * tasks:(i + 1) * tasks]) for i, k in enumerate(det._TARGET_KEYS)}
        out = det.center_head.loss(targets, det.center_head(det.neck(bev)))
        if self.keys is None:
            self.keys = sorted(out)
        return tuple(out[k] for k in self.keys)


def build_model(self, config, backend=None):
    return VoxelNet(config, backend=backend)
