"""CPU experiment for the next precision mode of the sparse-conv kernels: relative error of
a K-term dot product computed with split operands and fp32 accumulation, against fp64.
  tf32      : single-pass TF32 (operands truncated to 10 mantissa bits)
  tf32x3    : a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with tf32 hi / lo            (the current default, 3 kind::tf32 MMAs)
  bf16x3    : the same three products with bf16 hi / lo                       (3 kind::f16 MMAs at twice the rate)
  bf16x6    : hi/mid/lo bf16 terms, 6 products                                (fp32-faithful on the bf16 pipe)
Every product of two tf32 / bf16 values is exact in fp32, so an fp32 matmul of the split operands reproduces what
the tensor core accumulates (up to summation order)."""
import torch

torch.manual_seed(0)


def tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def split_tf32(x):
    hi = tf32_trunc(x)
    return hi, tf32_trunc(x - hi)


def split_bf16(x, terms):
    out, r = [], x
    for _ in range(terms):
        t = r.bfloat16().float()
        out.append(t)
        r = r - t
    return out


def report(m, k, n):
    a = torch.randn(m, k)
    b = torch.randn(k, n) * 0.05
    ref = a.double() @ b.double()
    scale = ref.abs().mean().item()
    res = {}
    res["fp32"] = a @ b
    res["tf32"] = tf32_trunc(a) @ tf32_trunc(b)
    ah, al = split_tf32(a)
    bh, bl = split_tf32(b)
    res["tf32x3"] = ah @ bh + al @ bh + ah @ bl
    a1, a2 = split_bf16(a, 2)
    b1, b2 = split_bf16(b, 2)
    res["bf16x3"] = a1 @ b1 + a2 @ b1 + a1 @ b2
    a1, a2, a3 = split_bf16(a, 3)
    b1, b2, b3 = split_bf16(b, 3)
    res["bf16x6"] = a1 @ b1 + a2 @ b1 + a1 @ b2 + a2 @ b2 + a3 @ b1 + a1 @ b3
    line = "K=%5d " % k
    for name, v in res.items():
        err = (v.double() - ref).abs()
        line += " %s max %.1e mean %.1e |" % (name, err.max().item() / scale, err.mean().item() / scale)
    print(line)


if __name__ == "__main__":
    print("errors relative to mean |result| (fp64 reference)")
    for k in (432, 1728, 3456, 6912):  # 27 taps x C_in for C_in = 16 .. 256
        report(512, k, 64)
