"""Find which part of the static section breaks CUDA-graph capture: capture each piece (forward only) on its own."""
import os, sys, traceback
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench

if "sidestream" in sys.argv:
    _main_stream = torch.cuda.Stream()
    torch.cuda.set_stream(_main_stream)   # nothing ever runs on the legacy default stream
class A: pass
args = A(); args.workload = "voxel_detr"
model, spec, cfg = bench.build_workload(args, "cuda:0")
model.train()
scenes = bench.make_scenes(2, 150000, seed=1, spec=spec)
batch = [({"points": torch.from_numpy(p).cuda()}, {"annotations": a}) for p, a in scenes]
losses = model(batch); bench.loss_total(losses).backward()
torch.cuda.synchronize()
if "dellosses" in sys.argv:
    del losses
    import gc; gc.collect()
det = model
with torch.no_grad():
    feats = det.bottom_up_maps(batch)
names = [n for n in det.backbone.extractor.in_features if n in feats]
maps = {n: feats[n].detach().clone() for n in names}

def try_capture(name, fn):
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    try:
        with torch.cuda.stream(s):
            for _ in range(2):
                out = fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay(); torch.cuda.synchronize()
        print("OK   ", name)
        return out
    except Exception as e:
        print("FAIL ", name, type(e).__name__, str(e).splitlines()[0][:200])
        traceback.print_exc(limit=6)
        torch.cuda.synchronize()
        return None

which = sys.argv[1:] or ["fpn", "proj", "encode", "proposals", "decoder", "heads", "fwdbwd_encoder_layer"]
with torch.no_grad():
    f = det.backbone.extractor.forward_dense(maps)
    p3 = f["p3"]
    pos = det.backbone.position_encoding(p3).type_as(p3)
    src = det.input_proj[0](p3)
    memory, anchors, shapes, start = det.transformer.encode([src], [pos])
    _, _, proposals, topk = det.transformer._get_enc_proposals(memory, anchors)
    hs, inter = det.transformer.decoder(None, None, memory, shapes, start, proposals)
    if "fpn" in which: try_capture("fpn.forward_dense", lambda: det.backbone.extractor.forward_dense(maps))
    if "proj" in which: try_capture("position_encoding + input_proj", lambda: det.input_proj[0](p3) + det.backbone.position_encoding(p3))
    if "encode" in which: try_capture("transformer.encode", lambda: det.transformer.encode([src], [pos]))
    if "proposals" in which: try_capture("proposals", lambda: det.transformer._get_enc_proposals(memory, anchors))
    if "decoder" in which: try_capture("decoder", lambda: det.transformer.decoder(None, None, memory, shapes, start, proposals))
    if "heads" in which: try_capture("heads", lambda: det.transformer.decoder.detection_head(hs[0], proposals[..., :7], 0))
if "fwdbwd_encoder_layer" in which:
    layer = det.transformer.encoder.layers[0]
    flat = src.flatten(2).transpose(1, 2).detach().clone().requires_grad_(True)
    fpos = pos.flatten(2).transpose(1, 2).detach()
    def fb():
        out = layer(flat, fpos, shapes, start, anchors)
        g, = torch.autograd.grad(out.sum(), flat)
        return g
    try_capture("encoder layer fwd+bwd", fb)

# ---- forward + backward of each piece under capture
def fb_of(name, fn, inputs, mods=()):
    if "params" in which:
        inputs = list(inputs) + [p for m in mods for p in m.parameters() if p.requires_grad]
    def run():
        outs = fn()
        outs = [o for o in (outs if isinstance(outs, (tuple, list)) else [outs]) if torch.is_tensor(o) and o.requires_grad]
        return torch.autograd.grad([o.sum() for o in outs], inputs, allow_unused=True)
    try_capture(name + " fwd+bwd", run)

if "bwd" in which or len(sys.argv) == 1:
    m2 = {n: v.clone().requires_grad_(True) for n, v in maps.items()}
    fb_of("fpn", lambda: det.backbone.extractor.forward_dense(m2)["p3"], list(m2.values()), [det.backbone.extractor.fpn_lateral3, det.backbone.extractor.fpn_output3, det.backbone.extractor.fpn_lateral4])
    p3g = p3.detach().clone().requires_grad_(True)
    fb_of("input_proj", lambda: det.input_proj[0](p3g), [p3g], [det.input_proj])
    srcg = src.detach().clone().requires_grad_(True)
    fb_of("encode", lambda: det.transformer.encode([srcg], [pos])[0], [srcg], [det.transformer.encoder])
    memg = memory.detach().clone().requires_grad_(True)
    fb_of("proposal head", lambda: det.transformer.proposal_head(memg, anchors), [memg], [det.transformer.proposal_head])
    fb_of("decoder", lambda: det.transformer.decoder(None, None, memg, shapes, start, proposals)[0], [memg], [det.transformer.decoder.layers])
    hsg = hs.detach().clone().requires_grad_(True)
    fb_of("heads", lambda: det.transformer.decoder.detection_head(hsg[0], proposals[..., :7], 0), [hsg], [det.transformer.decoder.detection_head])
    from efg_b200.detectors.voxel_detr.model import _StaticSection
    sec = _StaticSection(det, names)
    m3 = [v.clone().requires_grad_(True) for v in maps.values()]
    fb_of("whole static section", lambda: sec(*m3)[:2], m3)

if "mgc" in which:
    import traceback
    try:
        sec2 = _StaticSection(det, names) if "_StaticSection" in globals() else None
        from efg_b200.detectors.voxel_detr.model import _StaticSection as SS
        sec2 = SS(det, names)
        sample = tuple(v.clone().requires_grad_(True) for v in maps.values())
        torch.cuda.synchronize()
        call = torch.cuda.make_graphed_callables(sec2, sample, allow_unused_input=True)
        out = call(*sample)
        (out[0].sum() + out[1].sum()).backward()
        torch.cuda.synchronize()
        print("OK    make_graphed_callables")
    except Exception:
        traceback.print_exc()
