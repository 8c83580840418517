"""tf_ops/grouping/tf_grouping.py mirror: query_ball_point(radius, nsample, xyz1, xyz2) (:9-21),
group_point(points, idx) (:48-56) with its registered gradient (:57-61).  query_ball_point2 /
select_top_k / knn_point are unused by DH3D and not built."""
import torch

from .. import ops


def query_ball_point(radius, nsample, xyz1, xyz2):
    """xyz1 [B,n,3] dataset, xyz2 [B,m,3] queries -> (idx [B,m,nsample] i32, pts_cnt [B,m] i32)."""
    return ops.query_ball_point(radius, nsample, xyz1, xyz2)


class _GroupPointFn(torch.autograd.Function):
    """@tf.RegisterGradient('GroupPoint') (tf_grouping.py:57-61): [group_point_grad(points, idx, grad_out), None]."""

    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(points, idx)
        return ops.group_point(points.detach(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        points, idx = ctx.saved_tensors
        return ops.group_point_grad(points, idx, grad_out.contiguous()), None


def group_point(points, idx):
    """points [B,N,C], idx [B,M,S] i32 -> [B,M,S,C]."""
    if torch.is_grad_enabled() and points.requires_grad:
        return _GroupPointFn.apply(points, idx)
    return ops.group_point(points, idx)


def group_point_grad(points, idx, grad_out):
    """grad_out [B,M,S,C] -> grad_points [B,N,C] (tf_grouping_g.cu:114-133)."""
    return ops.group_point_grad(points, idx, grad_out)
