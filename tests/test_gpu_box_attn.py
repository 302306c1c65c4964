"""GPU parity: box attention forward/backward vs the golden vectors produced by the reference's
own torch twin (ms_deform_attn_core_pytorch) and vs the CPU oracle at Voxel-DETR geometry.
Tolerance 1e-4 absolute on fp32 values of O(1) (north_star: 1e-3)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import box_attn as obox

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
BOX_CASES = sorted(glob.glob(os.path.join(GOLDEN, "box_attn_*.pt")))


def _level_start(shapes):
    return torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])


@pytest.mark.parametrize("path", BOX_CASES, ids=[os.path.basename(p)[9:-3] for p in BOX_CASES])
def test_box_attn_matches_reference_golden(path):
    from efg_b200.operators import BoxAttnFunction

    g = torch.load(path)
    shapes = g["shapes"].cuda()
    v = g["value"].cuda().requires_grad_(True)
    l = g["loc"].cuda().requires_grad_(True)
    a = g["attn"].cuda().requires_grad_(True)
    out = BoxAttnFunction.apply(v, shapes, _level_start(shapes), l, a, 64)
    assert torch.allclose(out.detach().cpu(), g["out"], atol=1e-5)
    out.backward(g["grad_out"].cuda())
    assert torch.allclose(v.grad.cpu(), g["grad_value"], atol=1e-4)
    assert torch.allclose(l.grad.cpu(), g["grad_loc"], atol=1e-3, rtol=1e-4)
    assert torch.allclose(a.grad.cpu(), g["grad_attn"], atol=1e-4)


def test_box_attn_voxel_detr_geometry_vs_oracle():
    """One level 47x47 (quarter of the 188x188 BEV map), 8 heads x 32 ch, 25 points, attention
    passed as [B,LQ,H,L,5,5] exactly like Box3dAttention does (VD/modules/box_attention.py:108)."""
    from efg_b200.operators import BoxAttnFunction

    gen = torch.Generator().manual_seed(7)
    B, H, C, hh, ww, P = 2, 8, 32, 47, 47, 25
    LQ = hh * ww
    shapes = torch.tensor([[hh, ww]], dtype=torch.int64)
    value = torch.randn(B, hh * ww, H, C, generator=gen)
    ys, xs = torch.meshgrid(torch.linspace(0.5, hh - 0.5, hh) / hh, torch.linspace(0.5, ww - 0.5, ww) / ww,
                            indexing="ij")
    centre = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)[None, :, None, None, None, :]
    loc = centre + (torch.rand(B, LQ, H, 1, P, 2, generator=gen) - 0.5) * 0.08
    attn = torch.softmax(torch.randn(B, LQ, H, P, generator=gen), -1).view(B, LQ, H, 1, 5, 5)
    grad_out = torch.randn(B, LQ, H * C, generator=gen)
    eo, egv, egl, ega = obox.forward_backward(value, shapes, loc, attn.view(B, LQ, H, 1, P), grad_out)
    v = value.cuda().requires_grad_(True)
    l = loc.cuda().requires_grad_(True)
    a = attn.cuda().requires_grad_(True)
    out = BoxAttnFunction.apply(v, shapes.cuda(), _level_start(shapes).cuda(), l, a, 64)
    assert torch.allclose(out.detach().cpu(), eo, atol=1e-5)
    out.backward(grad_out.cuda())
    assert torch.allclose(v.grad.cpu(), egv, atol=2e-4)
    assert torch.allclose(l.grad.cpu(), egl, atol=2e-3, rtol=1e-4)
    assert torch.allclose(a.grad.cpu().view(B, LQ, H, 1, P), ega, atol=1e-4)


def test_box_attn_linearity_in_value_full_size():
    """Size-independent property at the full 188x188 map: the op is linear in `value` and in `attn`."""
    from efg_b200 import ops

    gen = torch.Generator(device="cuda").manual_seed(3)
    B, H, C, hh, ww, P, LQ = 1, 8, 32, 188, 188, 25, 4096
    shapes = torch.tensor([[hh, ww]], dtype=torch.int64, device="cuda")
    start = torch.zeros(1, dtype=torch.int64, device="cuda")
    v1 = torch.randn(B, hh * ww, H, C, device="cuda", generator=gen)
    v2 = torch.randn(B, hh * ww, H, C, device="cuda", generator=gen)
    loc = torch.rand(B, LQ, H, 1, P, 2, device="cuda", generator=gen) * 1.1 - 0.05
    attn = torch.rand(B, LQ, H, 1, P, device="cuda", generator=gen)
    o1 = ops.box_attn_forward(v1, shapes, start, loc, attn)
    o2 = ops.box_attn_forward(v2, shapes, start, loc, attn)
    o12 = ops.box_attn_forward(v1 + 2 * v2, shapes, start, loc, attn)
    assert torch.allclose(o12, o1 + 2 * o2, atol=1e-4)
    o3 = ops.box_attn_forward(v1, shapes, start, loc, attn * 3)
    assert torch.allclose(o3, o1 * 3, atol=1e-4)
    # all-outside sampling locations give exactly zero and zero gradients
    far = torch.full_like(loc, 7.0)
    assert ops.box_attn_forward(v1, shapes, start, far, attn).abs().max().item() == 0.0
    gv, gl, ga = ops.box_attn_backward(v1, shapes, start, far, attn, torch.ones(B, LQ, H * C, device="cuda"))
    assert gv.abs().max().item() == 0 and gl.abs().max().item() == 0 and ga.abs().max().item() == 0


def test_box_attn_im2col_contract():
    from efg_b200 import _C

    v = torch.randn(3, 16, 2, 8, device="cuda")
    shapes = torch.tensor([[4, 4]], device="cuda")
    start = torch.tensor([0], device="cuda")
    loc = torch.rand(3, 5, 2, 1, 4, 2, device="cuda")
    attn = torch.rand(3, 5, 2, 1, 4, device="cuda")
    with pytest.raises(RuntimeError):
        _C.box_attn_forward(v, shapes, start, loc, attn, 2)  # 3 % 2 != 0
    assert _C.box_attn_forward(v, shapes, start, loc, attn, 64).shape == (3, 5, 16)


@pytest.mark.parametrize("with_rotation", [False, True])
def test_fused_grid_softmax_matches_torch_chain(with_rotation):
    """The fused sampling-grid + softmax kernel vs the reference's chain of torch ops
    (Box3dAttention._where_to_attend + F.softmax), forward and gradients, 1e-5."""
    from efg_b200 import ops
    from efg_b200.detectors.voxel_detr.box_attention import Box3dAttention
    from oracle.backend_cpu import cpu_backend

    torch.manual_seed(4)
    B, LQ, d, H = 2, 700, 256, 8
    mod = Box3dAttention(d, 1, H, with_rotation=with_rotation, backend=cpu_backend()).cuda()  # torch-chain branch
    with torch.no_grad():
        mod.linear_box_weight.normal_(0, 0.2)
        mod.linear_attn_weight.normal_(0, 0.2)
    query = torch.randn(B, LQ, d, device="cuda")
    ref = torch.rand(B, LQ, 7, device="cuda")
    ref[..., 3:5] = ref[..., 3:5] * 0.1 + 0.01
    # reference chain
    q1 = query.clone().requires_grad_(True)
    attn1 = torch.softmax(torch.nn.functional.linear(q1, mod.linear_attn_weight, mod.linear_attn_bias).view(B, LQ, H, -1), -1)
    grid1 = mod._where_to_attend(q1, None, ref)
    g_grid, g_attn = torch.randn_like(grid1), torch.randn_like(attn1)
    (grid1 * g_grid).sum().add((attn1 * g_attn).sum()).backward()
    # fused
    q2 = query.clone().requires_grad_(True)
    offsets = torch.nn.functional.linear(q2, mod.linear_box_weight, mod.linear_box_bias).view(B, LQ, H, 1, mod.num_variable)
    logits = torch.nn.functional.linear(q2, mod.linear_attn_weight, mod.linear_attn_bias).view(B, LQ, H, -1)
    grid2, attn2 = ops.BoxGridSoftmaxFunction.apply(offsets, logits, ref, mod.kernel_indices)
    (grid2 * g_grid).sum().add((attn2 * g_attn).sum()).backward()
    assert grid2.shape == grid1.shape
    assert (grid2 - grid1).abs().max().item() < 1e-5
    assert (attn2 - attn1).abs().max().item() < 1e-6
    assert (q2.grad - q1.grad).abs().max().item() < 1e-4 * max(1.0, q1.grad.abs().max().item())


@pytest.mark.parametrize("with_rotation", [False, True])
def test_box3d_attention_merged_projection_matches_torch_chain(with_rotation):
    """Encoder-sized inputs take the merged path (both projections as ONE tensor-core GEMM feeding the strided
    grid/softmax kernel); it must agree with the reference's chain of separate F.linear + torch ops on the same
    parameters, forward and gradients."""
    from efg_b200.detectors.voxel_detr.box_attention import Box3dAttention
    from oracle.backend_cpu import cpu_backend

    torch.manual_seed(5)
    hh = ww = 72
    B, LQ, d, H = 1, hh * ww, 256, 8  # 5184 rows >= the 4096-row threshold of the dense tensor-core path
    ref_mod = Box3dAttention(d, 1, H, with_rotation=with_rotation, backend=cpu_backend()).cuda()
    fast_mod = Box3dAttention(d, 1, H, with_rotation=with_rotation).cuda()
    with torch.no_grad():
        ref_mod.linear_box_weight.normal_(0, 0.05)
        ref_mod.linear_attn_weight.normal_(0, 0.05)
    fast_mod.load_state_dict(ref_mod.state_dict())
    shapes = torch.tensor([[hh, ww]], dtype=torch.int64, device="cuda")
    start = torch.zeros(1, dtype=torch.int64, device="cuda")
    ys, xs = torch.meshgrid(torch.linspace(0.5, hh - 0.5, hh, device="cuda") / hh,
                            torch.linspace(0.5, ww - 0.5, ww, device="cuda") / ww, indexing="ij")
    ref = torch.zeros(B, LQ, 7, device="cuda")
    ref[..., 0], ref[..., 1] = xs.reshape(-1), ys.reshape(-1)
    ref[..., 3:5] = 0.06
    ref[..., 6] = 0.1
    query = torch.randn(B, LQ, d, device="cuda")
    value = torch.randn(B, LQ, d, device="cuda")
    g = torch.randn(B, LQ, d, device="cuda")
    outs = []
    for mod in (ref_mod, fast_mod):
        q = query.clone().requires_grad_(True)
        v = value.clone().requires_grad_(True)
        out, attn = mod(q, v, shapes, None, start, None, ref)
        (out * g).sum().backward()
        if attn is None:   # fused attention path: the weights are never materialised; compare the reference's with themselves
            attn = outs[0][1]
        outs.append((out.detach(), attn.detach(), q.grad, v.grad, mod.linear_box_weight.grad, mod.linear_attn_weight.grad,
                     mod.linear_box_bias.grad, mod.linear_attn_bias.grad))
    names = ("out", "attn", "dquery", "dvalue", "dW_box", "dW_attn", "db_box", "db_attn")
    for n, a, b in zip(names, *outs):
        scale = max(1.0, a.abs().max().item())
        diff = (a - b).abs()
        if n == "dquery":
            # d(bilinear sample)/d(location) is piecewise constant per BEV cell: a sampling location that lands within
            # 1 ulp of a cell border may floor() to the other cell in one of the two implementations, which changes
            # that query's gradient by O(1).  Allow a vanishing fraction of such rows, require the rest to agree.
            bad_rows = (diff.amax(-1) > 1e-3 * scale).float().mean().item()
            assert bad_rows < 5e-3, (n, bad_rows)
        elif n in ("dW_box", "db_box"):
            # ... and those few rows enter the box-projection gradients, which sum over all rows
            rel = (diff.norm() / a.norm().clamp_min(1e-12)).item()
            assert rel < 2e-2, (n, rel)
        else:
            # the merged projection runs on the tensor cores in the default bf16x3 mode (error ~2.5e-5 relative per GEMM);
            # the bar is north_star's 1e-3
            assert diff.max().item() < 1e-3 * scale, (n, diff.max().item(), scale)


@pytest.mark.parametrize("nv", [4, 5])
@pytest.mark.parametrize("hw,lq_is_grid", [((24, 28), True), ((20, 20), False)])
def test_fused_where_to_attend_equals_the_two_operator_chain(nv, hw, lq_is_grid):
    """BoxAttnProjFunction (sampling grid + softmax inside the attention kernels) against BoxProjGridSoftmaxFunction ->
    BoxAttnFunction on the same projection output: outputs equal, gradients of value and of the projection equal up to
    the order of the atomic accumulation."""
    from efg_b200 import ops
    from efg_b200.operators.box_attention_func import BoxAttnFunction

    torch.manual_seed(nv + hw[0])
    b, h, p = 2, 8, 25
    lv = hw[0] * hw[1]
    lq = lv if lq_is_grid else 300
    n_attn, n_box = h * p, h * nv
    ld = (n_attn + n_box + 63) // 64 * 64
    value = torch.randn(b, lv, h, 32, device="cuda")
    proj = torch.randn(b, lq, ld, device="cuda")
    ref = torch.rand(b, lq, 7, device="cuda")
    ref[..., 3:5] = ref[..., 3:5] * 0.2 + 0.02
    ref[..., 0:2] = ref[..., 0:2] * 0.9 + 0.05
    k = torch.linspace(-2, 2, 5)
    i, j = torch.meshgrid(k, k, indexing="ij")
    kidx = (torch.stack([j, i], dim=-1).view(-1, 2) / 5).cuda()
    shapes = torch.tensor([list(hw)], dtype=torch.int64, device="cuda")
    start = torch.zeros(1, dtype=torch.int64, device="cuda")
    go = torch.randn(b, lq, h * 32, device="cuda")

    v1, p1 = value.clone().requires_grad_(True), proj.clone().requires_grad_(True)
    loc, attn = ops.BoxProjGridSoftmaxFunction.apply(p1, ref, kidx, h, 1, nv)
    o1 = BoxAttnFunction.apply(v1, shapes, start, loc, attn.view(b, lq, h, 1, 5, 5), 64)
    o1.backward(go)
    v2, p2 = value.clone().requires_grad_(True), proj.clone().requires_grad_(True)
    o2 = ops.BoxAttnProjFunction.apply(v2, shapes, start, p2, ref, kidx, h, nv, hw[1] if lq_is_grid else 0)
    o2.backward(go)
    # same formulas in both paths; what differs is FMA contraction (a few ulp) and the order of the atomic accumulation
    assert torch.allclose(o2, o1, rtol=1e-5, atol=2e-5), float((o2 - o1).abs().max())
    assert torch.allclose(v2.grad, v1.grad, rtol=1e-4, atol=1e-4), float((v2.grad - v1.grad).abs().max())
    scale = max(float(p1.grad.abs().max()), 1.0)
    # d(bilinear sample)/d(location) is piecewise constant: a location within an ulp of a cell border may floor() to the
    # other cell in one of the two paths -> allow a vanishing fraction of rows to differ
    bad_rows = ((p2.grad - p1.grad).abs().amax(-1) > 1e-4 * scale).float().mean().item()
    assert bad_rows <= 1e-3, bad_rows
    if ld > n_attn + n_box:
        assert float(p2.grad[..., n_attn + n_box:].abs().max()) == 0.0
