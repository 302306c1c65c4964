"""Where does a Voxel-DETR step spend its time? CPU enqueue vs GPU busy, top CUDA kernels, sync count."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import make_scenes, NUM_QUERIES
from efg_b200.config import voxel_detr_config
from efg_b200.detectors.voxel_detr import VoxelDETR

dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = voxel_detr_config(model={"device": "cuda:0", "transformer": {"num_queries": NUM_QUERIES}})
model = VoxelDETR(cfg).train()
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
batch = [(torch.from_numpy(p).to(dev), a) for p, a in make_scenes(2, 150000, 1)]

def step():
    opt.zero_grad(set_to_none=True)
    losses = model([({"points": p}, {"annotations": a}) for p, a in batch])
    total = sum(v for k, v in losses.items() if k.startswith("loss"))
    total.backward()
    opt.step()
    return total

for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("enqueue %.1f ms  total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ka = prof.key_averages()
cuda_total = sum(e.self_device_time_total for e in ka) / 1e3
print("sum of CUDA kernel time: %.1f ms" % cuda_total)
print(ka.table(sort_by="self_device_time_total", row_limit=28, max_name_column_width=70))
syncs = [e for e in ka if "ynchronize" in e.key or "item" in e.key or "to_copy" in e.key or "_local_scalar" in e.key]
for e in syncs: print(e.key, e.count, "cpu ms %.2f" % (e.cpu_time_total / 1e3))
