"""tf_ops/interpolation/tf_interpolate.py mirror: three_nn(xyz1, xyz2) (:8-17),
three_interpolate(points, idx, weight) (:19-28).  CPU-only ops in the reference
(tf_interpolate.cpp:187,222); GPU kernels here."""
from .. import ops


def three_nn(xyz1, xyz2):
    """xyz1 [B,n,3] unknown, xyz2 [B,m,3] known -> (dist [B,n,3] SQUARED, idx [B,n,3] i32)."""
    return ops.three_nn(xyz1, xyz2)


def three_interpolate(points, idx, weight):
    """points [B,m,c], idx [B,n,3], weight [B,n,3] -> [B,n,c]."""
    return ops.three_interpolate(points, idx, weight)
