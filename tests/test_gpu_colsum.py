"""GPU parity: efgb_colsum (bias gradient of the token-wise linears) vs torch's fp64 column sum."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols", [(1, 4), (7, 16), (1000, 40), (4097, 256), (70688, 256), (70688, 1024), (33333, 1028)])
def test_colsum_matches_fp64(rows, cols):
    from efg_b200 import ops

    gen = torch.Generator(device="cuda").manual_seed(rows + cols)
    x = torch.randn(rows, cols, device="cuda", generator=gen)
    got = ops.colsum(x)
    want = x.double().sum(0)
    assert got.shape == (cols,)
    assert (got.double() - want).abs().max().item() < 2e-5 * max(1.0, rows ** 0.5)
    # deterministic: two calls are bit-identical
    assert torch.equal(got, ops.colsum(x))


def test_colsum_rejects_bad_inputs():
    from efg_b200 import ops

    with pytest.raises(RuntimeError):
        ops.colsum(torch.randn(8, 6, device="cuda"))  # cols % 4 != 0
    with pytest.raises(RuntimeError):
        ops.colsum(torch.randn(8, 8))  # CPU tensor
    assert ops.colsum(torch.zeros(0, 8, device="cuda")).abs().sum().item() == 0.0
