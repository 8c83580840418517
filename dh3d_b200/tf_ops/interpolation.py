"""tf_ops/interpolation/tf_interpolate.py mirror: three_nn(xyz1, xyz2) (:8-17),
three_interpolate(points, idx, weight) (:19-28) with its registered gradient (:29-34).  CPU-only ops
in the reference (tf_interpolate.cpp:187,222,264); GPU kernels here."""
import torch

from .. import ops


def three_nn(xyz1, xyz2):
    """xyz1 [B,n,3] unknown, xyz2 [B,m,3] known -> (dist [B,n,3] SQUARED, idx [B,n,3] i32)."""
    return ops.three_nn(xyz1, xyz2)


class _ThreeInterpolateFn(torch.autograd.Function):
    """@tf.RegisterGradient('ThreeInterpolate') (tf_interpolate.py:29-34): gradient to points only."""

    @staticmethod
    def forward(ctx, points, idx, weight):
        ctx.save_for_backward(points, idx, weight)
        return ops.three_interpolate(points.detach(), idx, weight.detach())

    @staticmethod
    def backward(ctx, grad_out):
        points, idx, weight = ctx.saved_tensors
        return ops.three_interpolate_grad(points, idx, weight.detach(), grad_out.contiguous()), None, None


def three_interpolate(points, idx, weight):
    """points [B,m,c], idx [B,n,3], weight [B,n,3] -> [B,n,c]."""
    if torch.is_grad_enabled() and points.requires_grad:
        return _ThreeInterpolateFn.apply(points, idx, weight)
    return ops.three_interpolate(points, idx, weight)


def three_interpolate_grad(points, idx, weight, grad_out):
    """grad_out [B,n,c] -> grad_points [B,m,c] (tf_interpolate.cpp:131-153)."""
    return ops.three_interpolate_grad(points, idx, weight, grad_out)
