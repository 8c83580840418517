"""tf_ops/grouping/tf_grouping.py mirror: query_ball_point(radius, nsample, xyz1, xyz2) (:9-21),
group_point(points, idx) (:48-56).  query_ball_point2 / select_top_k / knn_point are unused by
DH3D and not built."""
from .. import ops


def query_ball_point(radius, nsample, xyz1, xyz2):
    """xyz1 [B,n,3] dataset, xyz2 [B,m,3] queries -> (idx [B,m,nsample] i32, pts_cnt [B,m] i32)."""
    return ops.query_ball_point(radius, nsample, xyz1, xyz2)


def group_point(points, idx):
    """points [B,N,C], idx [B,M,S] i32 -> [B,M,S,C]."""
    return ops.group_point(points, idx)
