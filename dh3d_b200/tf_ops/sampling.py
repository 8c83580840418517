"""tf_ops/sampling/tf_sampling.py mirror: farthest_point_sample(npoint, inp) (:63-71, npoint
FIRST), gather_point(inp, idx) (:38-46).  prob_sample (:15-23) is unused by DH3D and not built."""
from .. import ops


def farthest_point_sample(npoint, inp):
    """inp [B,N,3] f32 -> [B,npoint] i32; starts at index 0; reference tie order (fps.cu)."""
    return ops.farthest_point_sample(npoint, inp)


def gather_point(inp, idx):
    """inp [B,N,3], idx [B,M] i32 -> [B,M,3]."""
    return ops.gather_point(inp, idx)
