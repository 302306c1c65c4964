import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
tf32 = len(sys.argv) > 1 and sys.argv[1] == "tf32"
torch.backends.cudnn.allow_tf32 = tf32
torch.backends.cuda.matmul.allow_tf32 = False
from efg_b200.detectors.voxel_detr import VoxelDETR
from oracle.backend_cpu import cpu_backend, voxelized_sample
from test_model_cpu import small_batch, small_config
torch.manual_seed(0)
cpu = VoxelDETR(small_config("cpu", 40), backend=cpu_backend())
gpu = VoxelDETR(small_config("cuda", 40))
gpu.load_state_dict(cpu.state_dict())
gpo = VoxelDETR(small_config("cuda", 40), backend=cpu_backend())
gpo.load_state_dict(cpu.state_dict()); gpo.train()
cpu.train(); gpu.train()
scenes = small_batch(2, 6000, seed=11)
cfg = cpu.config
from test_gpu_model import _ReplayMatcher
rm = _ReplayMatcher(); rm.install(cpu, True)
lc = cpu([(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in scenes])
rm.replay = list(rm.log); rm.install(gpu, False)
lg = gpu([({"points": p}, {"annotations": a}) for p, a in scenes])
for k in lc:
    print("%-16s cpu %.6f gpu %.6f" % (k, float(lc[k]), float(lg[k])))
tc = sum(v for k, v in lc.items() if k.startswith("loss")); tg = sum(v for k, v in lg.items() if k.startswith("loss"))
tc.backward(); tg.backward()
rm.replay = list(rm.log); rm.install(gpo, False)
lo = gpo([({"points": p}, {"annotations": a}) for p, a in scenes])
to = sum(v for k, v in lo.items() if k.startswith("loss")); to.backward()
po = dict(gpo.named_parameters())
print("=== GPU-oracle(torch ops on cuda) vs CPU-oracle")
for name in po:
    gc, gg = dict(cpu.named_parameters())[name].grad, po[name].grad
    if gc is None or gg is None: continue
    scale = float(gc.abs().max()); err = float((gg.cpu() - gc).abs().max())
    if err > 1e-3 * scale + 1e-7: print("%-70s rel %.3e" % (name, err / max(scale, 1e-12)))
print("=== efgb-cuda vs GPU-oracle")
for name in po:
    gc, gg = po[name].grad, dict(gpu.named_parameters())[name].grad
    if gc is None or gg is None: continue
    scale = float(gc.abs().max()); err = float((gg - gc).abs().max())
    if err > 1e-3 * scale + 1e-7: print("%-70s rel %.3e" % (name, err / max(scale, 1e-12)))
print("=== efgb-cuda vs CPU-oracle")
pc, pg = dict(cpu.named_parameters()), dict(gpu.named_parameters())
for name in pc:
    gc, gg = pc[name].grad, pg[name].grad
    if gc is None or gg is None:
        if (gc is None) != (gg is None): print("NONE MISMATCH", name)
        continue
    scale = float(gc.abs().max())
    err = float((gg.cpu() - gc).abs().max())
    if err > 1e-3 * scale + 1e-7:
        print("%-70s rel %.3e (scale %.3e)" % (name, err / max(scale, 1e-12), scale))
print("done tf32=%s" % tf32)
