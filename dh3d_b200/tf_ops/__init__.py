"""Drop-in mirrors of the reference's PointNet++ ``tf_ops`` python wrappers (boundary B)."""
from .sampling import farthest_point_sample, gather_point, gather_point_grad  # noqa: F401
from .grouping import group_point, group_point_grad, query_ball_point  # noqa: F401
from .interpolation import three_interpolate, three_interpolate_grad, three_nn  # noqa: F401
