"""ORACLE (test infrastructure) — sparse 3-D convolution on the CPU.

The arithmetic of this part of the path lives in a third-party dependency that is ABSENT from
/root/reference and from this image: **spconv** (PyPI ``spconv-cu11x``; the reference pins no
version — README.md:26-27 only says ``spconv_cu11{X}``, setup.cfg:17-30 does not list it,
efg/modeling/backbones/sparse_net.py:6-11 accepts 1.x or 2.x).  **Parity unpinned**: the reference
holds no test, golden vector or fixture at this boundary, and spconv cannot be run here.
What this module restates is spconv's published algorithm, anchored on the reference's call sites
(sparse_net.py:85-95, 125-147, 273-282, 405-426, 485-524):
  * SubMConv3d: outputs = inputs (same order); out[i] = sum_k W[k] in[site(c_i + k - K//2)]
  * SparseConv3d: out_dim = floor((in + 2p - k)/s) + 1; an output exists iff >= 1 input maps to it;
    out[o] = sum_k W[k] in[site(o*s - p + k)]; cross-correlation orientation (torch.nn.Conv3d)
  * weight layout [Cout, kD, kH, kW, Cin] (spconv 2.x)
  * output row order of SparseConv3d inside upstream spconv is an implementation detail that cannot
    be verified here; this oracle (and the CUDA path) emit ascending linear (b,z,y,x) order and
    parity on rulebooks is defined on the canonically sorted pair set.
Two independent restatements cross-check each other in tests/test_oracle.py:
  (1) sparse: sorted int64 keys + searchsorted rulebook, per-tap gather-mm-index_add;
  (2) dense: scatter to a dense grid, torch.nn.functional.conv3d, read back at the output sites.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _triple(v):
    return [int(x) for x in v] if isinstance(v, (list, tuple)) else [int(v)] * 3


def linear_key(coords, dhw):
    c = np.asarray(coords, dtype=np.int64)
    d, h, w = [int(x) for x in dhw]
    return ((c[:, 0] * d + c[:, 1]) * h + c[:, 2]) * w + c[:, 3]


def tap_offsets(ksize):
    kd, kh, kw = _triple(ksize)
    return [(kz, ky, kx) for kz in range(kd) for ky in range(kh) for kx in range(kw)]


def subm_rulebook(coords, batch, dhw, ksize):
    """nbr [M, K] int64: row of the site at c_i + k - K//2, or -1."""
    coords = np.asarray(coords, dtype=np.int64)
    m = coords.shape[0]
    kd, kh, kw = _triple(ksize)
    taps = tap_offsets(ksize)
    nbr = np.full((m, len(taps)), -1, dtype=np.int64)
    if m == 0:
        return nbr
    keys = linear_key(coords, dhw)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    d, h, w = [int(x) for x in dhw]
    for t, (kz, ky, kx) in enumerate(taps):
        q = coords.copy()
        q[:, 1] += kz - kd // 2
        q[:, 2] += ky - kh // 2
        q[:, 3] += kx - kw // 2
        inb = (q[:, 1] >= 0) & (q[:, 1] < d) & (q[:, 2] >= 0) & (q[:, 2] < h) & (q[:, 3] >= 0) & (q[:, 3] < w)
        qk = linear_key(q, dhw)
        pos = np.searchsorted(skeys, qk)
        pos = np.clip(pos, 0, m - 1)
        hit = inb & (skeys[pos] == qk)
        nbr[hit, t] = order[pos[hit]]
    return nbr


def conv_out_shape(in_dhw, ksize, stride, padding):
    return [(int(i) + 2 * p - k) // s + 1 for i, k, s, p in zip(in_dhw, _triple(ksize), _triple(stride), _triple(padding))]


def sparse_rulebook(coords, batch, in_dhw, ksize, stride, padding):
    """-> out_coords [Mo,4] (ascending linear order), out_dhw, nbr [Mo,K], nbr_t [Mi,K]."""
    coords = np.asarray(coords, dtype=np.int64)
    k, s, p = _triple(ksize), _triple(stride), _triple(padding)
    out_dhw = conv_out_shape(in_dhw, k, s, p)
    taps = tap_offsets(k)
    m_in = coords.shape[0]
    cand_keys, cand_valid, cand_out = [], [], []
    for (kz, ky, kx) in taps:
        tz = coords[:, 1] + p[0] - kz
        ty = coords[:, 2] + p[1] - ky
        tx = coords[:, 3] + p[2] - kx
        ok = (tz >= 0) & (ty >= 0) & (tx >= 0) & (tz % s[0] == 0) & (ty % s[1] == 0) & (tx % s[2] == 0)
        oz, oy, ox = tz // s[0], ty // s[1], tx // s[2]
        ok &= (oz < out_dhw[0]) & (oy < out_dhw[1]) & (ox < out_dhw[2])
        oc = np.stack([coords[:, 0], oz, oy, ox], 1)
        cand_valid.append(ok)
        cand_out.append(oc)
        cand_keys.append(linear_key(oc, out_dhw))
    allk = np.concatenate([ck[v] for ck, v in zip(cand_keys, cand_valid)]) if m_in else np.zeros((0,), np.int64)
    ukeys = np.unique(allk)
    m_out = ukeys.shape[0]
    d, h, w = out_dhw
    out_coords = np.zeros((m_out, 4), dtype=np.int32)
    out_coords[:, 3] = ukeys % w
    out_coords[:, 2] = (ukeys // w) % h
    out_coords[:, 1] = (ukeys // (w * h)) % d
    out_coords[:, 0] = ukeys // (w * h * d)
    nbr = np.full((m_out, len(taps)), -1, dtype=np.int64)
    nbr_t = np.full((m_in, len(taps)), -1, dtype=np.int64)
    for t in range(len(taps)):
        v = cand_valid[t]
        rows_in = np.nonzero(v)[0]
        rows_out = np.searchsorted(ukeys, cand_keys[t][v])
        nbr[rows_out, t] = rows_in
        nbr_t[rows_in, t] = rows_out
    return out_coords, out_dhw, nbr, nbr_t


def canonical_pairs(nbr, in_coords, out_coords):
    """Order-independent rulebook: sorted array of (tap, in b,z,y,x, out b,z,y,x)."""
    nbr = np.asarray(nbr)
    o, k = np.nonzero(nbr >= 0)
    j = nbr[o, k]
    rows = np.concatenate([k[:, None], np.asarray(in_coords)[j], np.asarray(out_coords)[o]], axis=1).astype(np.int64)
    return rows[np.lexsort(rows.T[::-1])]


def conv(feats, weight, bias, nbr):
    """feats [Mi,Cin] torch f32, weight [Cout,kd,kh,kw,Cin] (spconv 2.x layout), nbr [Mo,K] -> [Mo,Cout].
    Per-tap gather - mm - index_add (spconv's 'native' algorithm); differentiable through torch."""
    nbr_t = torch.as_tensor(np.asarray(nbr), dtype=torch.long).to(feats.device)
    c_out = weight.shape[0]
    taps = nbr_t.shape[1]
    w = weight.reshape(c_out, taps, -1)
    out = feats.new_zeros((nbr_t.shape[0], c_out))
    for t in range(taps):
        o = torch.nonzero(nbr_t[:, t] >= 0, as_tuple=True)[0]
        if o.numel() == 0:
            continue
        out = out.index_add(0, o, feats[nbr_t[o, t]] @ w[:, t, :].t())
    if bias is not None:
        out = out + bias
    return out


def dense_conv_at_sites(feats, in_coords, batch, in_dhw, weight, bias, stride, padding, out_coords):
    """Restatement (2): dense conv3d over the scattered grid, read back at out_coords.
    NB: a dense conv adds the bias everywhere and is only equal to the sparse conv AT ACTIVE OUTPUT
    SITES; for SubM the outputs are the inputs, for SparseConv3d every site with >=1 contributing input."""
    c_in = feats.shape[1]
    d, h, w = [int(x) for x in in_dhw]
    dense = feats.new_zeros((batch, c_in, d, h, w))
    ic = torch.as_tensor(np.asarray(in_coords), dtype=torch.long)
    dense[ic[:, 0], :, ic[:, 1], ic[:, 2], ic[:, 3]] = feats
    wt = weight.permute(0, 4, 1, 2, 3).contiguous()  # [Cout, Cin, kd, kh, kw]
    y = F.conv3d(dense, wt, bias, stride=_triple(stride), padding=_triple(padding))
    oc = torch.as_tensor(np.asarray(out_coords), dtype=torch.long)
    return y[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]]


def to_dense(feats, coords, batch, dhw):
    d, h, w = [int(x) for x in dhw]
    dense = feats.new_zeros((batch, feats.shape[1], d, h, w))
    c = torch.as_tensor(np.asarray(coords), dtype=torch.long).to(feats.device)
    dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = feats
    return dense
