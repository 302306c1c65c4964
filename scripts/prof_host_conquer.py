"""Host profile of one ConQueR step in the bench configuration (encoder section graphed, prepare() pipeline)."""
import os, sys, time, argparse
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from efg_b200.parallel import GradAverager
from torch.profiler import profile, ProfilerActivity

args = argparse.Namespace(workload="conquer", scenes=2, points=150000)
dev = torch.device("cuda:0")
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
torch.manual_seed(0)
model, spec, cfg = bench.build_workload(args, "cuda:0")
model.train()
averager = GradAverager(model)
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, fused=True)
batches = [[(torch.from_numpy(p).to(dev), a) for p, a in bench.make_scenes(2, 150000, seed=1 + b, spec=spec)] for b in range(2)]
inputs_of = lambda b: [({"points": p}, {"annotations": a}) for p, a in b]
acc = {}
def wrap(obj, name, label):
    fn = getattr(obj, name)
    def timed(*a, **k):
        t0 = time.perf_counter(); out = fn(*a, **k)
        acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return out
    setattr(obj, name, timed)
def step(b):
    t0 = time.perf_counter()
    averager.zero_grad()
    losses = model(inputs_of(b)); t1 = time.perf_counter()
    total = bench.loss_total(losses); total.backward(); t2 = time.perf_counter()
    averager.finish(); averager.hide_unused(); opt.step(); t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
for i in range(3): step(batches[i % 2])
print("graph:", model.enable_static_graph(inputs_of(batches[0])), model.static_graph_error)
for i in range(3): step(batches[i % 2])
torch.cuda.synchronize()
wrap(model, "encode_targets", "encode_targets")
wrap(model, "bottom_up_maps", "bottom_up")
wrap(model, "conquer_losses", "conquer_losses")
wrap(model.transformer, "forward", "transformer.forward (decoders)")
wrap(model.transformer.decoder, "forward", "  decoder")
wrap(model.transformer.decoder_gt, "forward", "  decoder_gt")
import efg_b200.detectors.conquer.model as cm
wrap(cm, "prepare_for_cdn", "prepare_for_cdn")
N = 6
tot = [0, 0, 0]
for i in range(N):
    r = step(batches[i % 2]); torch.cuda.synchronize()
    tot = [a + b for a, b in zip(tot, r)]
print("host ms/step (synchronised between steps): forward %.1f backward %.1f optimizer %.1f" % tuple(t / N for t in tot))
for k, v in acc.items(): print("%-40s %7.2f ms" % (k, v / N))
with profile(activities=[ProfilerActivity.CPU]) as prof:
    step(batches[0]); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=55))
