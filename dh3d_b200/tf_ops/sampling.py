"""tf_ops/sampling/tf_sampling.py mirror: farthest_point_sample(npoint, inp) (:63-71, npoint
FIRST), gather_point(inp, idx) (:38-46) with its registered gradient (:47-51).  prob_sample (:15-23)
is unused by DH3D and not built."""
import torch

from .. import ops


def farthest_point_sample(npoint, inp):
    """inp [B,N,3] f32 -> [B,npoint] i32; starts at index 0; reference tie order (fps.cu)."""
    return ops.farthest_point_sample(npoint, inp)


class _GatherPointFn(torch.autograd.Function):
    """@tf.RegisterGradient('GatherPoint') (tf_sampling.py:47-51): [gather_point_grad(inp, idx, out_g), None]."""

    @staticmethod
    def forward(ctx, inp, idx):
        ctx.save_for_backward(inp, idx)
        return ops.gather_point(inp.detach(), idx)

    @staticmethod
    def backward(ctx, out_g):
        inp, idx = ctx.saved_tensors
        return ops.gather_point_grad(inp, idx, out_g.contiguous()), None


def gather_point(inp, idx):
    """inp [B,N,3], idx [B,M] i32 -> [B,M,3]."""
    if torch.is_grad_enabled() and inp.requires_grad:
        return _GatherPointFn.apply(inp, idx)
    return ops.gather_point(inp, idx)


def gather_point_grad(inp, idx, out_g):
    """out_g [B,M,3] -> inp_g [B,N,3] (tf_sampling_g.cu:183-192)."""
    return ops.gather_point_grad(inp, idx, out_g)
