"""Voxel-DETR transformer (VD/transformer.py:9-238): a box-attention encoder over the flattened
BEV map, top-k proposal selection from a 1-class head on the encoder output, and a decoder of
(dense self-attention over the queries -> box cross-attention into the BEV memory -> FFN) layers
with iterative reference-window refinement."""
import copy

import torch
import torch.nn.functional as F
from torch import nn

from ...backend import cuda_backend
from ...modeling.norm import TokenLinear
from .box_attention import Box3dAttention


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        # TokenLinear == nn.Linear (same parameters); with many rows (the proposal head runs on all 70k BEV
        # tokens) CUDA inputs go through the tcgen05 dense path, small inputs stay on cuBLAS
        self.layers = nn.ModuleList(TokenLinear(a, b, backend=cuda_backend()) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i < self.num_layers - 1:
                x = F.relu(x)
        return x


def get_clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


def _act(name):
    return {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}[name]


def _add_pos(t, pos):
    return t if pos is None else t + pos


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, nlevel, dim_feedforward, dropout, activation, backend=None):
        super().__init__()
        self.self_attn = Box3dAttention(d_model, nlevel, nhead, with_rotation=False, backend=backend)
        self.linear1 = TokenLinear(d_model, dim_feedforward, backend=backend or cuda_backend())
        self.dropout = nn.Dropout(dropout)
        self.linear2 = TokenLinear(dim_feedforward, d_model, backend=backend or cuda_backend())
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.activation = _act(activation)

    def forward(self, src, pos, src_shape, src_start_idx, ref_windows):
        attended = self.self_attn(_add_pos(src, pos), src, src_shape, None, src_start_idx, None, ref_windows)[0]
        src = self._add_norm(src, self.dropout1(attended), self.norm1)
        rows = src.numel() // src.shape[-1]
        if self._fast(src) and self.activation is F.relu and not (self.training and self.dropout.p > 0):
            from ... import ops

            if ops.fused_ffn_supported(rows, self.linear1.in_features, self.linear1.out_features, self.linear2.out_features):
                ffn = ops.fused_ffn(src, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)
                return self._add_norm(src, self.dropout2(ffn), self.norm2)
        ffn = self.linear2(self.dropout(self.activation(self.linear1(src))))
        return self._add_norm(src, self.dropout2(ffn), self.norm2)

    def _fast(self, t):
        return getattr(self.linear1, "_tc", False) and t.is_cuda and t.dtype == torch.float32

    def _add_norm(self, x, r, norm):
        """norm(x + r): one fused kernel each way on the CUDA backend (csrc/layernorm.cu), else the reference's two ops."""
        if self._fast(x) and norm.elementwise_affine and norm.bias is not None:
            from ... import ops

            if ops.add_layer_norm_supported(x.numel() // x.shape[-1], x.shape[-1]):
                return ops.add_layer_norm(x, r, norm.weight, norm.bias, norm.eps)
        return norm(x + r)


class TransformerEncoder(nn.Module):
    def __init__(self, d_model, encoder_layer, num_layers):
        super().__init__()
        self.layers = get_clones(encoder_layer, num_layers)

    def forward(self, src, pos, src_shape, src_start_idx, ref_windows):
        for layer in self.layers:
            src = layer(src, pos, src_shape, src_start_idx, ref_windows)
        return src


class TransformerDecoderLayer(nn.Module):
    def __init__(self, d_model, nhead, nlevel, dim_feedforward, dropout, activation, backend=None):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = Box3dAttention(d_model, nlevel, nhead, with_rotation=True, backend=backend)
        self.pos_embed_layer = MLP(10, d_model, d_model, 3)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.activation = _act(activation)

    def forward(self, idx, query, query_pos, memory, memory_shape, memory_start_idx, ref_windows, attn_mask=None):
        if idx == 0:
            # the first layer's content query IS the embedding of its reference window
            query = self.pos_embed_layer(ref_windows)
            q = k = query
        elif query_pos is None:
            query_pos = self.pos_embed_layer(ref_windows)
            q = k = _add_pos(query, query_pos)
        else:
            q = k = _add_pos(query, query_pos)
        sa = self.self_attn(q.transpose(0, 1), k.transpose(0, 1), query.transpose(0, 1), attn_mask=attn_mask)[0]
        query = self.norm1(query + self.dropout1(sa.transpose(0, 1)))
        ca = self.multihead_attn(_add_pos(query, query_pos), memory, memory_shape, None, memory_start_idx, None,
                                 ref_windows[..., :7])[0]
        query = self.norm2(query + self.dropout2(ca))
        ffn = self.linear2(self.dropout(self.activation(self.linear1(query))))
        return self.norm3(query + self.dropout3(ffn))


class TransformerDecoder(nn.Module):
    def __init__(self, d_model, decoder_layer, num_layers):
        super().__init__()
        self.layers = get_clones(decoder_layer, num_layers)
        self.detection_head = None  # attached by the detector

    def forward(self, query, query_pos, memory, memory_shape, memory_start_idx, ref_windows, attn_mask=None):
        output = query
        hidden, refs = [], []
        for idx, layer in enumerate(self.layers):
            output = layer(idx, output, query_pos, memory, memory_shape, memory_start_idx, ref_windows, attn_mask)
            logits, new_windows = self.detection_head(output, ref_windows[..., :7], idx)
            ref_windows = torch.cat((new_windows.detach(), logits.sigmoid().detach()), dim=-1)
            hidden.append(output)
            refs.append(new_windows)
        return torch.stack(hidden), torch.stack(refs)


class Transformer(nn.Module):
    def __init__(self, d_model=256, nhead=8, nlevel=4, num_encoder_layers=6, num_decoder_layers=6,
                 dim_feedforward=1024, dropout=0.1, activation="relu", num_queries=300, backend=None):
        super().__init__()
        self.num_queries = num_queries
        enc = TransformerEncoderLayer(d_model, nhead, nlevel, dim_feedforward, dropout, activation, backend)
        self.encoder = TransformerEncoder(d_model, enc, num_encoder_layers)
        dec = TransformerDecoderLayer(d_model, nhead, nlevel, dim_feedforward, dropout, activation, backend)
        self.decoder = TransformerDecoder(d_model, dec, num_decoder_layers)
        self.proposal_head = None  # attached by the detector
        self._ref_cache = {}
        self._enc_head_out = None  # (logits, windows) of the proposal head, reused by the encoder loss

    def _create_ref_windows(self, tensor_list):
        """One (cx, cy, 0.5, 0.025, 0.025, 0.5, 0) window per BEV cell (VD/transformer.py:30-52)."""
        out = []
        for t in tensor_list:
            B, _, H, W = t.shape
            key = (H, W, t.device)
            if key not in self._ref_cache:
                ys = torch.linspace(0.5, H - 0.5, H, dtype=torch.float32, device=t.device)
                xs = torch.linspace(0.5, W - 0.5, W, dtype=torch.float32, device=t.device)
                ry, rx = torch.meshgrid(ys, xs, indexing="ij")
                xy = torch.stack((rx.reshape(-1) / W, ry.reshape(-1) / H), -1)
                z = torch.zeros_like(xy[:, :1])
                self._ref_cache[key] = torch.cat((xy, z + 0.5, torch.ones_like(xy) * 0.025, z + 0.5, z), -1)
            out.append(self._ref_cache[key][None].expand(B, -1, -1))
        return torch.cat(out, dim=1)

    def _get_enc_proposals(self, enc_embed, ref_windows):
        logits, windows = self.proposal_head(enc_embed, ref_windows)
        self._enc_head_out = (logits, windows) if self.training and torch.is_grad_enabled() else None
        probs = logits[..., 0].sigmoid()
        # The reference takes topk(sorted=False): which of several EQUAL scores survive, and in which order,
        # is unspecified (empty BEV cells produce bit-identical scores).  Canonical choice here: a stable
        # descending sort (ties -> lowest BEV index), then ascending index order, so every device
        # enumerates the same query set identically.
        order = torch.argsort(probs, dim=1, descending=True, stable=True)[:, :self.num_queries]
        indexes, _ = order.sort(dim=1)
        topk_probs = torch.gather(probs, 1, indexes)
        indexes = indexes.unsqueeze(-1)
        windows = torch.gather(windows, 1, indexes.expand(-1, -1, windows.shape[-1]))
        windows = torch.cat((windows.detach(), topk_probs.detach().unsqueeze(-1).expand(-1, -1, 3)), dim=-1)
        return None, None, windows, indexes

    def encode(self, src, pos):
        assert pos is not None, "position encoding is required!"
        anchors = self._create_ref_windows(src)
        key = ("shapes", tuple((t.shape[2], t.shape[3]) for t in src), src[0].device)
        if key not in self._ref_cache:  # level shapes / start offsets are constants of the BEV geometry
            shapes = torch.tensor([[t.shape[2], t.shape[3]] for t in src], dtype=torch.int64, device=src[0].device)
            start = torch.cat([shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]])
            self._ref_cache[key] = (shapes, start)
        shapes, start = self._ref_cache[key]
        flat = torch.cat([t.flatten(2).transpose(1, 2) for t in src], dim=1)
        flat_pos = torch.cat([p.flatten(2).transpose(1, 2) for p in pos], dim=1)
        memory = self.encoder(flat, flat_pos, shapes, start, anchors)
        return memory, anchors, shapes, start

    def forward(self, src, pos):
        memory, anchors, shapes, start = self.encode(src, pos)
        query_embed, query_pos, proposals, topk_indexes = self._get_enc_proposals(memory, anchors)
        hs, inter_refs = self.decoder(query_embed, query_pos, memory, shapes, start, proposals)
        return hs, proposals[..., :7], inter_refs, memory, anchors, topk_indexes
