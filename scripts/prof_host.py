"""Host (Python + launch) time per phase of a Voxel-DETR step in the bench configuration (CUDA-graphed static section,
side stream), no synchronisation inside the step: which phases make the step host-bound."""
import os, sys, time, argparse
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from efg_b200 import ops
from efg_b200.parallel import GradAverager

args = argparse.Namespace(workload="voxel_detr", scenes=2, points=150000)
dev = torch.device("cuda:0")
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
torch.manual_seed(0)
model, spec, cfg = bench.build_workload(args, "cuda:0")
model.train()
averager = GradAverager(model)
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.01, betas=(0.9, 0.99), eps=1e-9)
batches = [[(torch.from_numpy(p).to(dev), a) for p, a in bench.make_scenes(2, 150000, seed=1 + b, spec=spec)] for b in range(2)]
acc = {}

def wrap(obj, name, label):
    fn = getattr(obj, name)
    def timed(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return out
    setattr(obj, name, timed)

side = torch.cuda.Stream(priority=-1)
def inputs_of(b): return [({"points": p}, {"annotations": a}) for p, a in b]
state = {"prep": None}
def step(b, rec=False, nxt=None):
    t = [time.perf_counter()]
    averager.zero_grad(); t.append(time.perf_counter())
    prep = state["prep"] or model.prepare(inputs_of(b), side)
    losses = model(inputs_of(b), prepared=prep); t.append(time.perf_counter())
    total = bench.loss_total(losses); t.append(time.perf_counter())
    total.backward(); t.append(time.perf_counter())
    averager.finish(); averager.hide_unused(); t.append(time.perf_counter())
    opt.step(); t.append(time.perf_counter())
    t0p = time.perf_counter()
    state["prep"] = model.prepare(inputs_of(nxt), side) if nxt is not None else None
    if rec: acc["prepare(next)"] = acc.get("prepare(next)", 0.0) + (time.perf_counter() - t0p) * 1e3
    if rec:
        for k, (a0, a1) in zip(["zero_grad", "model.forward", "loss sum", "backward", "averager", "optimizer"], zip(t[:-1], t[1:])):
            acc[k] = acc.get(k, 0.0) + (a1 - a0) * 1e3

for i in range(3): step(batches[i % 2])
torch.cuda.synchronize()
if not os.environ.get("NO_GRAPH"):
    print("graph:", model.enable_static_graph([({"points": p}, {"annotations": a}) for p, a in batches[0]]), model.static_graph_error)
for i in range(3): step(batches[i % 2])
torch.cuda.synchronize()
wrap(model, "encode_targets", "  fwd: encode_targets")
wrap(model, "bottom_up_maps", "  fwd: bottom_up (voxelize + sparse backbone)")
wrap(model, "losses", "  fwd: losses")
wrap(ops, "refresh_packs", "  fwd: refresh_packs")
if getattr(model, "_static_call", None) is not None:
    inner = model._static_call
    def timed_call(*a, **k):
        t0 = time.perf_counter()
        out = inner(*a, **k)
        acc["  fwd: static graph call"] = acc.get("  fwd: static graph call", 0.0) + (time.perf_counter() - t0) * 1e3
        return out
    object.__setattr__(model, "_static_call", timed_call) if not isinstance(inner, torch.nn.Module) else model.__dict__.__setitem__("_static_call", timed_call)
N = 10
t0 = time.perf_counter()
for i in range(N): step(batches[i % 2], True, batches[(i + 1) % 2])
state["prep"] = None
host = (time.perf_counter() - t0) * 1e3 / N
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / N
print("host %.2f ms/step, wall %.2f ms/step" % (host, wall))
for k, v in acc.items(): print("%-50s %8.2f ms" % (k, v / N))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU]) as prof:
    step(batches[0]); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=45, max_name_column_width=60))

# GPU busy fraction in the pipelined steady state
for i in range(4): step(batches[i % 2], False, batches[(i + 1) % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(4): step(batches[i % 2], False, batches[(i + 1) % 2])
    torch.cuda.synchronize()
state["prep"] = None
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ivs = sorted((e.time_range.start, e.time_range.end) for e in evs)
span = ivs[-1][1] - ivs[0][0]
busy, cur_s, cur_e = 0, ivs[0][0], ivs[0][1]
gaps = []
for s_, e_ in ivs[1:]:
    if s_ > cur_e:
        busy += cur_e - cur_s
        gaps.append((s_ - cur_e, cur_e))
        cur_s, cur_e = s_, e_
    else:
        cur_e = max(cur_e, e_)
busy += cur_e - cur_s
print("4 steps: span %.2f ms, GPU busy (union of kernels) %.2f ms = %.1f %%, sum of kernel time %.2f ms; idle %.2f ms in %d gaps" % (
    span / 1e3, busy / 1e3, 100.0 * busy / span, sum(e - s for s, e in ivs) / 1e3, (span - busy) / 1e3, len(gaps)))
gaps.sort(reverse=True)
print("largest gaps (us):", [round(g[0], 1) for g in gaps[:20]])
print("gaps > 20 us: %d totalling %.2f ms; gaps 5-20 us: %d totalling %.2f ms; gaps < 5 us: %d totalling %.2f ms" % (
    sum(1 for g in gaps if g[0] > 20), sum(g[0] for g in gaps if g[0] > 20) / 1e3,
    sum(1 for g in gaps if 5 < g[0] <= 20), sum(g[0] for g in gaps if 5 < g[0] <= 20) / 1e3,
    sum(1 for g in gaps if g[0] <= 5), sum(g[0] for g in gaps if g[0] <= 5) / 1e3))

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof2:
    step(batches[0], False, None); torch.cuda.synchronize()
ka = prof2.key_averages()
rows = sorted(ka, key=lambda e: -e.self_device_time_total)
print("top ops by self CUDA time (one step):")
tot = sum(e.self_device_time_total for e in rows)
for e in rows[:70]:
    if e.self_device_time_total < 50: break
    print("%-90s n=%5d  %8.3f ms  %5.1f %%" % (e.key[:90], e.count, e.self_device_time_total / 1e3, 100.0 * e.self_device_time_total / tot))
