"""Micro-benchmark of the box-attention kernels at the Voxel-DETR encoder geometry (B=2, one 188x188 level,
8 heads x 32 channels, 25 taps, queries = BEV cells, boxes as at initialisation) and at the decoder geometry
(300 queries).  EFGB_BOX_ATTN=generic selects the generic kernels for comparison."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from efg_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, H, C, hh, ww, P = 2, 8, 32, 188, 188, 25
shapes = torch.tensor([[hh, ww]], dtype=torch.int64, device=dev)
start = torch.zeros(1, dtype=torch.int64, device=dev)
value = torch.randn(B, hh * ww, H, C, device=dev)
kidx = torch.stack(torch.meshgrid(torch.arange(-2, 3.), torch.arange(-2, 3.), indexing="ij")[::-1], -1).view(-1, 2).to(dev) / 5


def make(lq, grid):
    if grid:
        ys, xs = torch.meshgrid(torch.linspace(0.5, hh - 0.5, hh, device=dev) / hh, torch.linspace(0.5, ww - 0.5, ww, device=dev) / ww, indexing="ij")
        centre = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)[None, :, None, None, None, :].expand(B, -1, H, 1, 1, 2)
    else:
        centre = torch.rand(B, lq, H, 1, 1, 2, device=dev)
    size = 0.025 * (1 + torch.rand(B, lq, H, 1, 1, 2, device=dev) / 8)
    loc = (centre + kidx.view(1, 1, 1, 1, P, 2) * size + torch.rand(B, lq, H, 1, 1, 2, device=dev) * 0.025 / 8).contiguous()
    attn = torch.softmax(torch.randn(B, lq, H, 1, P, device=dev), -1)
    go = torch.randn(B, lq, H * C, device=dev)
    return loc, attn, go


def timeit(fn, iters=int(os.environ.get("ITERS", "20")), warm=int(os.environ.get("WARM", "3"))):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / max(iters, 1) * 1e3


print("mode=%s" % os.environ.get("EFGB_BOX_ATTN", "tile"))
print("case           LQ     kind      us   alg GB/s")
for name, lq, grid in (("encoder", hh * ww, True), ("decoder", 300, False)):
    loc, attn, go = make(lq, grid)
    fb = 4 * (value.numel() + loc.numel() + attn.numel() + go.numel())
    bb = 4 * (2 * value.numel() + 2 * loc.numel() + 2 * attn.numel() + go.numel())
    us = timeit(lambda: ops.box_attn_forward(value, shapes, start, loc, attn))
    print("%-10s %7d  fwd  %9.1f  %8.1f" % (name, lq, us, fb / us / 1e3))
    us = timeit(lambda: ops.box_attn_backward(value, shapes, start, loc, attn, go))
    print("%-10s %7d  bwd  %9.1f  %8.1f" % (name, lq, us, bb / us / 1e3))

# the fused operator (sampling grid + softmax inside the attention kernels) on a projection output that yields the same
# boxes as above: logits random, offsets ~ U[0, 1) like linear_box_bias at initialisation
from efg_b200 import _lib
import ctypes
L = _lib.lib()
for nv in (4,):
    lq = hh * ww
    n_attn, n_box = H * P, H * nv
    ld = (n_attn + n_box + 63) // 64 * 64
    proj = torch.zeros(B, lq, ld, device=dev)
    proj[..., :n_attn] = torch.randn(B, lq, n_attn, device=dev)
    proj[..., n_attn:n_attn + n_box] = torch.rand(B, lq, n_box, device=dev)
    ys, xs = torch.meshgrid(torch.linspace(0.5, hh - 0.5, hh, device=dev) / hh, torch.linspace(0.5, ww - 0.5, ww, device=dev) / ww, indexing="ij")
    ref = torch.zeros(B, lq, 7, device=dev)
    ref[..., 0], ref[..., 1] = xs.reshape(-1), ys.reshape(-1)
    ref[..., 2] = ref[..., 5] = 0.5
    ref[..., 3:5] = 0.025
    out = torch.empty(B, lq, H * C, device=dev)
    go = torch.randn(B, lq, H * C, device=dev)
    gv, gp = torch.empty_like(value), torch.empty_like(proj)
    off = ctypes.c_void_p(proj.data_ptr() + 4 * n_attn)
    goff = ctypes.c_void_p(gp.data_ptr() + 4 * n_attn)
    st = ops._stream()
    fwd = lambda: L.efgb_box_attn_fused_forward(ops._p(value), ops._p(shapes), ops._p(start), off, ops._p(proj), ops._p(ref), ops._p(kidx),
                                                B, hh * ww, H, lq, P, nv, ld, ld, ww, ops._p(out), st)
    bwd = lambda: L.efgb_box_attn_fused_backward(ops._p(value), ops._p(shapes), ops._p(start), off, ops._p(proj), ops._p(ref), ops._p(kidx),
                                                 ops._p(go), B, hh * ww, H, lq, P, nv, ld, ld, ww, ops._p(gv), goff, ops._p(gp), st)
    fb = 4 * (value.numel() + B * lq * (n_attn + n_box) + out.numel())
    bb = 4 * (2 * value.numel() + 2 * B * lq * (n_attn + n_box) + go.numel())
    us = timeit(fwd)
    print("%-10s %7d  fwd  %9.1f  %8.1f   [fused grid + softmax]" % ("encoder", lq, us, fb / us / 1e3))
    us = timeit(bwd)
    print("%-10s %7d  bwd  %9.1f  %8.1f   [fused grid + softmax]" % ("encoder", lq, us, bb / us / 1e3))
