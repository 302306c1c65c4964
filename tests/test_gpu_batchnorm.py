"""GPU: the fused BatchNorm1d (+ residual) (+ ReLU) pass (csrc/batchnorm.cu, ops.bn_act) against torch's own
nn.BatchNorm1d / add / relu chain in fp64 — outputs, running statistics, and every gradient; the operand planes it can
emit equal ops.split_bf16 of its output; SparseSequential takes the fused path and stays equal to the CPU oracle."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols", [(1, 16), (37, 16), (4099, 32), (30000, 64), (9000, 128), (700, 256)])
@pytest.mark.parametrize("relu,with_res", [(False, False), (True, False), (True, True)])
def test_bn_act_training_matches_torch_fp64(rows, cols, relu, with_res):
    from efg_b200 import ops

    torch.manual_seed(rows + cols)
    x = (torch.randn(rows, cols, device="cuda") * 2 + 0.5).requires_grad_(True)
    res = torch.randn(rows, cols, device="cuda", requires_grad=True) if with_res else None
    bn = nn.BatchNorm1d(cols).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    ref = nn.BatchNorm1d(cols).cuda().double().train()
    ref.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in bn.state_dict().items()})
    xd = x.detach().double().requires_grad_(True)
    rd = res.detach().double().requires_grad_(True) if with_res else None
    if rows == 1:
        with pytest.raises(Exception):   # torch refuses batch statistics over a single row; the kernel does not crash
            ref(xd)
        y = ops.bn_act(x, bn, residual=res, relu=relu)
        assert torch.isfinite(y).all()
        return
    yd = ref(xd)
    if with_res:
        yd = yd + rd
    if relu:
        yd = torch.relu(yd)
    y = ops.bn_act(x, bn, residual=res, relu=relu)
    g = torch.randn_like(y)
    y.backward(g)
    yd.backward(g.double())
    assert (y.double() - yd).abs().max().item() < 2e-5
    assert (bn.running_mean.double() - ref.running_mean).abs().max().item() < 1e-6
    assert (bn.running_var.double() - ref.running_var).abs().max().item() < 1e-5
    assert int(bn.num_batches_tracked) == 1
    scale = max(1.0, xd.grad.abs().max().item())
    assert (x.grad.double() - xd.grad).abs().max().item() < 2e-5 * scale
    assert (bn.weight.grad.double() - ref.weight.grad).abs().max().item() < 1e-4 * max(1.0, ref.weight.grad.abs().max().item())
    assert (bn.bias.grad.double() - ref.bias.grad).abs().max().item() < 1e-4 * max(1.0, ref.bias.grad.abs().max().item())
    if with_res:
        assert (res.grad.double() - rd.grad).abs().max().item() < 1e-6


def test_bn_act_eval_mode_and_planes():
    from efg_b200 import ops

    torch.manual_seed(1)
    x = torch.randn(5000, 64, device="cuda")
    bn = nn.BatchNorm1d(64).cuda()
    with torch.no_grad():
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2.0)
    bn.eval()
    y, planes = ops.bn_act(x, bn, relu=True, want_planes=True)
    assert torch.allclose(y, torch.relu(bn(x)), atol=1e-5)
    assert torch.equal(planes, ops.split_bf16(y))


def test_sparse_sequential_uses_the_fused_pass_and_matches_the_oracle():
    import numpy as np
    from efg_b200 import ops, spconv
    from oracle import spconv_cpu

    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    dhw, m = [9, 40, 41], 3000
    cells = np.sort(rng.choice(2 * dhw[0] * dhw[1] * dhw[2], size=m, replace=False))
    coords = np.stack([cells // (dhw[0] * dhw[1] * dhw[2]), (cells // (dhw[1] * dhw[2])) % dhw[0], (cells // dhw[2]) % dhw[1],
                       cells % dhw[2]], 1).astype(np.int32)
    feats = torch.randn(m, 16)

    def build(sp):
        return sp.SparseSequential(sp.SubMConv3d(16, 32, 3, padding=1, bias=False, indice_key="a"), nn.BatchNorm1d(32), nn.ReLU(),
                                   sp.SubMConv3d(32, 32, 3, padding=1, bias=False, indice_key="a"), nn.BatchNorm1d(32))

    cpu_seq, gpu_seq = build(spconv_cpu), build(spconv)
    gpu_seq.load_state_dict(cpu_seq.state_dict())
    gpu_seq.cuda()
    calls = []
    orig = ops.bn_act
    ops.bn_act = lambda *a, **k: (calls.append(k.get("want_planes", False)), orig(*a, **k))[1]
    try:
        fg = feats.cuda().requires_grad_(True)
        shortcut = spconv.SparseConvTensor(torch.randn(m, 32, device="cuda"), torch.from_numpy(coords).cuda(), dhw, 2)
        yg = gpu_seq(spconv.SparseConvTensor(fg, torch.from_numpy(coords).cuda(), dhw, 2), residual=shortcut, final_relu=True)
    finally:
        ops.bn_act = orig
    assert calls == [True, False]   # first norm feeds a tensor-core conv (planes emitted), the last one ends the block
    fc = feats.clone().requires_grad_(True)
    yc = cpu_seq(spconv_cpu.SparseConvTensor(fc, torch.from_numpy(coords), dhw, 2))
    yc = torch.relu(yc.features + shortcut.features.cpu())
    assert (yg.features.detach().cpu() - yc.detach()).abs().max().item() < 1e-3
    yg.features.square().sum().backward()
    yc.square().sum().backward()
    assert (fg.grad.cpu() - fc.grad).abs().max().item() < 2e-3 * max(1.0, fc.grad.abs().max().item())
