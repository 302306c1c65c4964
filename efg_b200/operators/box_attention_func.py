"""``BoxAttnFunction`` — mirror of efg/operators/box_attention_func.py:9-64."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _C


class BoxAttnFunction(Function):
    """apply(value, value_spatial_shapes, value_level_start_index, sampling_locations,
    attention_weights, im2col_step) -> [B, LQ, H*C].  Inputs are computed in fp32 (the reference
    wraps forward in ``custom_fwd(cast_inputs=torch.float32)``)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        value = value.contiguous()
        sampling_locations = sampling_locations.contiguous()
        attention_weights = attention_weights.contiguous()
        output = _C.box_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                     attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_contiguous():
            grad_output = grad_output.contiguous()
        value, shapes, level_start, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = _C.box_attn_backward(
            value, shapes, level_start, sampling_locations, attention_weights, grad_output, ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None
