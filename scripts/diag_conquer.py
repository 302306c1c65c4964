import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
torch.backends.cudnn.allow_tf32 = False
from efg_b200.config import conquer_config
from efg_b200.detectors.conquer import ConQueR
from efg_b200.detectors.conquer.cdn import draw_noise
from oracle.backend_cpu import cpu_backend, voxelized_sample
from test_model_cpu import SMALL, small_batch
from test_gpu_model import _ReplayMatcher
def cfg(device):
    return conquer_config(dataset={"pc_range": SMALL.pc_range, "voxel_size": SMALL.voxel_size, "max_voxel_num": 20000},
                          model={"device": device, "transformer": {"num_queries": 40, "enc_layers": 1, "dec_layers": 2}})
torch.manual_seed(0)
cpu = ConQueR(cfg("cpu"), backend=cpu_backend()).train()
gpu = ConQueR(cfg("cuda")).train()
gpu.load_state_dict(cpu.state_dict())
scenes = small_batch(2, 6000, seed=21)
total_gt = sum(len(a["labels"]) for _, a in scenes)
noise = draw_noise(2 * 3 * total_gt, 3, torch.device("cpu"), generator=torch.Generator().manual_seed(5))
cpu.cdn_noise = gpu.cdn_noise = noise
bc = [(voxelized_sample(p, cpu.config.dataset), {"annotations": a}) for p, a in scenes]
bg = [({"points": p}, {"annotations": a}) for p, a in scenes]
with torch.no_grad():
    fc, pc = cpu.extract(bc); fg, pg = gpu.extract(bg)
    print("feat diff", (fg[0].cpu() - fc[0]).abs().max().item())
    mc = cpu.transformer.encode(fc, pc); mg = gpu.transformer.encode(fg, pg)
    print("memory diff", (mg[0].cpu() - mc[0]).abs().max().item())
    _, _, wc, ic = cpu.transformer._get_enc_proposals(mc[0], mc[1]); _, _, wg, ig = gpu.transformer._get_enc_proposals(mg[0], mg[1])
    print("topk equal", torch.equal(ic, ig.cpu()), (ic != ig.cpu()).sum().item())
    lc_, _ = cpu.transformer.proposal_head(mc[0], mc[1]); lg_, _ = gpu.transformer.proposal_head(mg[0], mg[1])
    pcs = lc_[..., 0].sigmoid(); s, _ = pcs.sort(dim=1, descending=True)
    print("score gap around k", (s[:, 38:42]).tolist(), "max logit diff", (lg_.cpu() - lc_).abs().max().item())
rm = _ReplayMatcher(); rm.install(cpu, True)
lc = cpu(bc); rm.replay = list(rm.log); rm.install(gpu, False); lg = gpu(bg)
for k in lc: print("%-28s %.6f %.6f" % (k, float(lc[k]), float(lg[k])))
