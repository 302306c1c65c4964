"""GPU: CenterPoint label assignment on the device (assign_batch_device: vectorised per-object scalars + one kernel that
draws every Gaussian of the batch) against the reference's host loop (assign_scene, CP/voxelnet.py:44-192) — indices,
masks, classes and regression targets exactly, heatmaps to float32 rounding; single- and multi-task heads, objects
outside the map, unknown class names, more objects than max_objs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WAYMO_TASKS = [{"num_classes": 3, "class_names": ["VEHICLE", "PEDESTRIAN", "CYCLIST"]}]
NUSC_TASKS = [{"num_classes": 1, "class_names": ["car"]}, {"num_classes": 2, "class_names": ["truck", "construction_vehicle"]},
              {"num_classes": 2, "class_names": ["bus", "trailer"]}, {"num_classes": 1, "class_names": ["barrier"]},
              {"num_classes": 2, "class_names": ["motorcycle", "bicycle"]}, {"num_classes": 2, "class_names": ["pedestrian", "traffic_cone"]}]


def _scene(rng, n, names, span):
    boxes = np.zeros((n, 9), np.float32)
    boxes[:, :2] = rng.uniform(-span * 1.1, span * 1.1, (n, 2))     # some centres fall outside the map
    boxes[:, 2] = rng.uniform(-1, 1, n)
    boxes[:, 3:6] = rng.uniform(0.3, 6.0, (n, 3))
    boxes[:, 6:8] = rng.normal(0, 2, (n, 2))
    boxes[:, 8] = rng.uniform(-7, 7, n)
    return {"annotations": {"gt_boxes": boxes, "gt_names": rng.choice(names + ["UNKNOWN"], n)}}


@pytest.mark.parametrize("tasks,pc_range,voxel,max_objs", [
    (WAYMO_TASKS, [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], [0.1, 0.1, 0.15], 500),
    (NUSC_TASKS, [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], [0.075, 0.075, 0.2], 500),
    (WAYMO_TASKS, [-12.8, -12.8, -2.0, 12.8, 12.8, 4.0], [0.1, 0.1, 0.15], 20),   # more objects than slots
])
def test_device_assignment_equals_host_loop(tasks, pc_range, voxel, max_objs):
    from efg_b200.detectors.centerpoint.assign import assign_batch_device, assign_scene

    rng = np.random.default_rng(len(tasks) + max_objs)
    names = [n for t in tasks for n in t["class_names"]]
    grid = np.round((np.asarray(pc_range[3:], np.float32) - np.asarray(pc_range[:3], np.float32)) / np.asarray(voxel, np.float32)).astype(np.int64)
    infos = [_scene(rng, n, names, pc_range[3]) for n in (70, 0, 45)]
    dev = assign_batch_device(infos, tasks, grid, pc_range, voxel, 8, 0.1, max_objs, 2, "cuda")
    host = [assign_scene(i["annotations"], tasks, grid, pc_range, voxel, 8, 0.1, max_objs, 2) for i in infos]
    for t in range(len(tasks)):
        for key in ("ind", "mask", "cat"):
            exp = np.stack([h[key][t] for h in host])
            assert np.array_equal(dev[key][t].cpu().numpy().astype(exp.dtype), exp), (t, key)
        exp = np.stack([h["anno_box"][t] for h in host])
        assert np.array_equal(dev["anno_box"][t].cpu().numpy(), exp), t
        exp = np.stack([h["hm"][t] for h in host])
        got = dev["hm"][t].cpu().numpy()
        assert got.shape == exp.shape and np.abs(got - exp).max() < 1e-6, (t, np.abs(got - exp).max())
        assert np.array_equal(got == 1.0, exp == 1.0)   # the peaks
