"""Box3dAttention (VD/modules/box_attention.py:10-115): every query predicts, per head and level,
a box (offsets w.r.t. its reference window, optionally a rotation) and attends to a 5x5 grid of
bilinear samples inside it, weighted by a softmax over the grid.

Parameter names match the reference (linear_box_weight/bias, linear_attn_weight/bias,
value_proj, out_proj, buffer kernel_indices).  The sampling itself is the box-attention CUDA
kernel (csrc/box_attn.cu) through the backend's ``box_attn`` callable."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from ...backend import cuda_backend
from ...modeling.norm import TokenLinear


class Box3dAttention(nn.Module):
    def __init__(self, d_model, num_level, num_head, with_rotation=True, kernel_size=5, backend=None):
        super().__init__()
        assert d_model % num_head == 0, "d_model should be divided by num_head"
        self._backend = [backend or cuda_backend()]
        self.im2col_step = 64
        self.d_model = d_model
        self.num_head = num_head
        self.num_level = num_level
        self.head_dim = d_model // num_head
        self.with_rotation = with_rotation
        self.num_variable = 5 if with_rotation else 4
        self.kernel_size = kernel_size
        self.num_point = kernel_size ** 2

        self.linear_box_weight = nn.Parameter(torch.zeros(num_level * num_head * self.num_variable, d_model))
        self.linear_box_bias = nn.Parameter(torch.zeros(num_head * num_level * self.num_variable))
        self.linear_attn_weight = nn.Parameter(torch.zeros(num_head * num_level * self.num_point, d_model))
        self.linear_attn_bias = nn.Parameter(torch.zeros(num_head * num_level * self.num_point))
        self.value_proj = TokenLinear(d_model, d_model, backend=self._backend[0])
        self.out_proj = TokenLinear(d_model, d_model, backend=self._backend[0])

        # grid of kernel offsets in units of the box size: (x, y) pairs, x fastest
        if kernel_size % 2 == 0:
            idx = torch.linspace(-kernel_size // 2 + 0.5, kernel_size // 2 - 0.5, kernel_size)
        else:
            idx = torch.linspace(-(kernel_size - 1) // 2, (kernel_size - 1) // 2, kernel_size)
        i, j = torch.meshgrid(idx, idx, indexing="ij")
        self.register_buffer("kernel_indices", torch.stack([j, i], dim=-1).view(-1, 2) / kernel_size)
        self._reset_parameters()

    def _reset_parameters(self):
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.out_proj.bias, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.0)
        nn.init.constant_(self.linear_attn_weight, 0.0)
        nn.init.constant_(self.linear_attn_bias, 0.0)
        nn.init.constant_(self.linear_box_weight, 0.0)
        nn.init.uniform_(self.linear_box_bias)

    def _where_to_attend(self, query, v_valid_ratios, ref_windows):
        """-> sampling grid [B, L, H, levels, P, 2] in normalised (x, y)."""
        B, L = ref_windows.shape[:2]
        offset = F.linear(query, self.linear_box_weight, self.linear_box_bias)
        offset = offset.view(B, L, self.num_head, self.num_level, self.num_variable)
        ref = ref_windows.unsqueeze(2).unsqueeze(3) if ref_windows.dim() == 3 else ref_windows.unsqueeze(3)
        ref_boxes = torch.cat((ref[..., 0:2], ref[..., 3:5]), dim=-1)  # (cx, cy, w, l); slices, no index upload
        ref_angles = ref[..., 6:7]
        if self.with_rotation:
            offset, offset_angles = offset.split(4, dim=-1)
            angles = (ref_angles + offset_angles / 16) * 2 * math.pi
        else:
            angles = ref_angles.expand(B, L, self.num_head, self.num_level, 1)
        size = ref_boxes[..., 2:4]
        boxes = ref_boxes + offset / 8 * torch.cat((size, size), dim=-1)
        center, size = boxes.unsqueeze(-2).split(2, dim=-1)
        cos, sin = torch.cos(angles), torch.sin(angles)
        rot = torch.stack([cos, -sin, sin, cos], dim=-1).view(B, L, self.num_head, self.num_level, 1, 2, 2)
        grid = self.kernel_indices * torch.relu(size)
        grid = center + (grid.unsqueeze(-2) * rot).sum(-1)
        if v_valid_ratios is not None:
            grid = grid * v_valid_ratios
        return grid.contiguous()

    def forward(self, query, value, v_shape, v_mask, v_start_index, v_valid_ratios, ref_windows):
        B, LQ = query.shape[:2]
        LV = value.shape[1]
        value = self.value_proj(value)
        if v_mask is not None:
            value = value.masked_fill(v_mask[..., None], float(0))
        value = value.view(B, LV, self.num_head, self.head_dim)
        fused = (self._backend[0].name == "efgb200-cuda" and query.is_cuda and v_valid_ratios is None and
                 ref_windows.dim() == 3 and ref_windows.shape[-1] == 7)
        attn = None
        if fused:
            # fused sampling grid + softmax (csrc/box_attn.cu), same math as the torch chain in the else-branch
            from ... import ops

            n_attn, n_box = self.linear_attn_weight.shape[0], self.linear_box_weight.shape[0]
            ld = (n_attn + n_box + 63) // 64 * 64  # 256 for 8 heads x 25 taps (+ 4|5 box variables)
            if ops.dense_linear_supported(B * LQ, self.d_model, ld):
                # many rows (encoder): both projections as ONE tensor-core GEMM [attn logits | box offsets | 0-pad]
                pad = ld - n_attn - n_box
                ws = [self.linear_attn_weight, self.linear_box_weight]
                bs = [self.linear_attn_bias, self.linear_box_bias]
                if pad:
                    ws.append(self.linear_attn_weight.new_zeros((pad, self.d_model)))
                    bs.append(self.linear_attn_bias.new_zeros((pad,)))
                proj = ops.dense_linear(query, torch.cat(ws, 0), torch.cat(bs, 0))
                if ops.box_attn_fused_supported(self.head_dim, self.num_level, self.num_point, self.num_variable):
                    # sampling grid + softmax inside the attention kernels: `grid` / `attn` never exist in memory (the
                    # attention weights, which no caller of this module uses, are not returned on this path)
                    out = ops.BoxAttnProjFunction.apply(value, v_shape, v_start_index, proj, ref_windows, self.kernel_indices,
                                                        self.num_head, self.num_variable,
                                                        ops._query_grid_width(v_shape, self.num_level, LQ))
                    return self.out_proj(out), None
                grid, attn = ops.BoxProjGridSoftmaxFunction.apply(proj, ref_windows, self.kernel_indices, self.num_head,
                                                                  self.num_level, self.num_variable)
            else:
                attn = F.linear(query, self.linear_attn_weight, self.linear_attn_bias)
                offsets = F.linear(query, self.linear_box_weight, self.linear_box_bias)
                grid, attn = ops.BoxGridSoftmaxFunction.apply(
                    offsets.view(B, LQ, self.num_head, self.num_level, self.num_variable),
                    attn.view(B, LQ, self.num_head, -1), ref_windows, self.kernel_indices)
            attn = attn.view(B, LQ, self.num_head, self.num_level, self.kernel_size, self.kernel_size)
        else:
            attn = F.linear(query, self.linear_attn_weight, self.linear_attn_bias)
            attn = F.softmax(attn.view(B, LQ, self.num_head, -1), dim=-1)
            attn = attn.view(B, LQ, self.num_head, self.num_level, self.kernel_size, self.kernel_size)
            grid = self._where_to_attend(query, v_valid_ratios, ref_windows)
        out = self._backend[0].box_attn(value, v_shape, v_start_index, grid, attn, self.im2col_step)
        return self.out_proj(out), attn
