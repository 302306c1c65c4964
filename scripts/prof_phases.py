"""Per-phase CPU-enqueue vs wall time of one Voxel-DETR step (a synchronize at every phase boundary):
a phase whose enqueue time ~ its wall time is launch-bound (CPU), one whose wall time is much larger is GPU-bound."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import make_scenes, NUM_QUERIES
from efg_b200.config import voxel_detr_config
from efg_b200.detectors.voxel_detr import VoxelDETR

dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = voxel_detr_config(model={"device": "cuda:0", "transformer": {"num_queries": NUM_QUERIES}})
model = VoxelDETR(cfg).train()
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
batch = [(torch.from_numpy(p).to(dev), a) for p, a in make_scenes(2, 150000, 1)]
rows = []

def wrap(obj, name, label):
    fn = getattr(obj, name)
    def timed(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        rows.append((label, (t1 - t0) * 1e3, (t2 - t0) * 1e3))
        return out
    setattr(obj, name, timed)

def step(record):
    opt.zero_grad(set_to_none=True)
    losses = model([({"points": p}, {"annotations": a}) for p, a in batch])
    total = sum(v for k, v in losses.items() if k.startswith("loss"))
    if record:
        torch.cuda.synchronize(); t0 = time.perf_counter()
    total.backward()
    if record:
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        rows.append(("backward", (t1 - t0) * 1e3, (t2 - t0) * 1e3)); t0 = time.perf_counter()
    opt.step()
    if record:
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        rows.append(("optimizer", (t1 - t0) * 1e3, (t2 - t0) * 1e3))

for _ in range(3): step(False)
wrap(model, "encode_targets", "encode_targets")
wrap(model, "voxelize_on_device", "  voxelize")
wrap(model.backbone, "forward", "  backbone(sparse+fpn)")
wrap(model, "extract", "extract(total)")
wrap(model.transformer, "encode", "  encoder")
wrap(model.transformer, "_get_enc_proposals", "  proposals(topk)")
wrap(model.transformer.decoder, "forward", "  decoder")
wrap(model.transformer, "forward", "transformer(total)")
wrap(model, "losses", "losses(total)")
wrap(model.transformer.decoder.detection_head.losses.matcher, "solve", "  hungarian(host)")
for it in range(2):
    rows.clear(); step(True)
print("%-26s %10s %10s" % ("phase", "enqueue ms", "wall ms"))
for label, enq, wall in rows: print("%-26s %10.2f %10.2f" % (label, enq, wall))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU]) as prof:
    step(False); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
