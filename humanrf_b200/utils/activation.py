"""truncated_exp, mirroring humanrf/utils/activation.py:6-39 (exp forward in fp32, gradient
exp(clamp(x, -15, 15))).  The fused kernels implement the same pair internally
(csrc/field_fwd.cu, csrc/field_bwd.cu); this torch version serves callers outside the kernels."""
import torch


class _truncated_exp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, threshold):
        x = x.float()
        ctx.save_for_backward(x)
        ctx.threshold = threshold
        return torch.exp(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return dy * torch.exp(x.clamp(-ctx.threshold, ctx.threshold)), None


def truncated_exp(inp: torch.Tensor, threshold: float = 15) -> torch.Tensor:
    return _truncated_exp.apply(inp, threshold)
