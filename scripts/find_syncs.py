import os, sys, warnings, collections, traceback
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import make_scenes, NUM_QUERIES
from efg_b200.config import voxel_detr_config
from efg_b200.detectors.voxel_detr import VoxelDETR
dev = torch.device("cuda:0")
cfg = voxel_detr_config(model={"device": "cuda:0", "transformer": {"num_queries": NUM_QUERIES}})
model = VoxelDETR(cfg).train()
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
batch = [(torch.from_numpy(p).to(dev), a) for p, a in make_scenes(2, 150000, 1)]
def step():
    opt.zero_grad(set_to_none=True)
    losses = model([({"points": p}, {"annotations": a}) for p, a in batch])
    total = sum(v for k, v in losses.items() if k.startswith("loss"))
    total.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
counts = collections.Counter()
def showwarning(message, category, filename, lineno, file=None, line=None):
    st = traceback.extract_stack()
    frames = [f for f in st if "/root/repo" in f.filename or "efg_b200" in f.filename]
    key = " <- ".join("%s:%d" % (os.path.basename(f.filename), f.lineno) for f in frames[-3:])
    if not frames:
        key = "EXT " + " <- ".join("%s:%d" % (os.path.basename(f.filename), f.lineno) for f in st[-6:-2])
    counts[key] += 1
warnings.showwarning = showwarning
warnings.simplefilter("always")
torch.cuda.set_sync_debug_mode("warn")
step()
torch.cuda.set_sync_debug_mode("default")
for k, v in counts.most_common(40): print(v, k)
