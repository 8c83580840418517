"""Drop-in mirrors of the reference's PointNet++ ``tf_ops`` python wrappers (boundary B)."""
from .sampling import farthest_point_sample, gather_point  # noqa: F401
from .grouping import group_point, query_ball_point  # noqa: F401
from .interpolation import three_interpolate, three_nn  # noqa: F401
