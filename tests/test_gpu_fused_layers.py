"""GPU parity of the fused encoder-layer pieces against the reference's chain of torch ops
(VD/transformer.py:56-64: norm(src + x), linear2(relu(linear1(src)))) — forward and gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_library_math():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("rows,cols", [(4096, 256), (5000, 128), (70688, 256), (4099, 512)])
def test_add_layer_norm_matches_torch(rows, cols):
    from efg_b200 import ops

    gen = torch.Generator(device="cuda").manual_seed(rows + cols)
    x = torch.randn(rows, cols, device="cuda", generator=gen) * 2 + 0.3
    r = torch.randn(rows, cols, device="cuda", generator=gen)
    w = torch.rand(cols, device="cuda", generator=gen) + 0.5
    b = torch.randn(cols, device="cuda", generator=gen)
    g = torch.randn(rows, cols, device="cuda", generator=gen)
    outs = []
    for fused in (False, True):
        xs, rs, ws, bs = (t.clone().requires_grad_(True) for t in (x, r, w, b))
        y = ops.add_layer_norm(xs, rs, ws, bs, 1e-5) if fused else F.layer_norm(xs + rs, (cols,), ws, bs, 1e-5)
        (y * g).sum().backward()
        outs.append((y.detach(), xs.grad, rs.grad, ws.grad, bs.grad))
    for name, a, c in zip(("y", "dx", "dr", "dgamma", "dbeta"), *outs):
        scale = max(1.0, a.abs().max().item())
        assert (a - c).abs().max().item() < 2e-5 * scale * (rows ** 0.5 if name in ("dgamma", "dbeta") else 1.0), name
    assert ops.add_layer_norm_supported(rows, cols)
    assert not ops.add_layer_norm_supported(rows, 200)


@pytest.mark.parametrize("rows,d,dff", [(4096, 256, 1024), (9001, 256, 512)])
def test_fused_ffn_matches_torch_chain(rows, d, dff):
    from efg_b200 import ops

    gen = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, d, device="cuda", generator=gen)
    w1 = torch.randn(dff, d, device="cuda", generator=gen) * 0.06
    b1 = torch.randn(dff, device="cuda", generator=gen) * 0.1
    w2 = torch.randn(d, dff, device="cuda", generator=gen) * 0.03
    b2 = torch.randn(d, device="cuda", generator=gen) * 0.1
    g = torch.randn(rows, d, device="cuda", generator=gen)
    assert ops.fused_ffn_supported(rows, d, dff, d)
    outs = []
    for fused in (False, True):
        ts = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
        y = ops.fused_ffn(*ts) if fused else F.linear(F.relu(F.linear(ts[0], ts[1], ts[2])), ts[3], ts[4])
        (y * g).sum().backward()
        outs.append([y.detach()] + [t.grad for t in ts])
    # ReLU has a kink: a pre-activation within rounding distance of 0 can be cut differently by two correct GEMMs
    # (cuBLAS fp32 vs the 3xTF32 kernel), which changes that element's gradient by O(1).  Values are compared
    # elementwise, gradients in the relative Frobenius norm plus an outlier bound on the rows of dx.
    y0, y1 = outs[0][0], outs[1][0]
    assert (y0 - y1).abs().max().item() < 2e-4 * max(1.0, y0.abs().max().item())
    for name, a, c in zip(("dx", "dw1", "db1", "dw2", "db2"), outs[0][1:], outs[1][1:]):
        rel = ((a - c).norm() / a.norm().clamp_min(1e-12)).item()
        assert rel < 2e-3, (name, rel)
    dx0, dx1 = outs[0][1], outs[1][1]
    bad_rows = ((dx0 - dx1).abs().amax(-1) > 2e-4 * max(1.0, dx0.abs().max().item())).float().mean().item()
    assert bad_rows < 1e-2, bad_rows
