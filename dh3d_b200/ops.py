"""Native (point-major) operator layer: thin torch-tensor wrappers over the C ABI.

Layouts here are the library's native ones -- features [B,N,C], neighbourhoods [B,N,K], xyz
[B,N,3] -- and every op can fuse the reference's follow-up BatchNorm/bias/activation.  The
reference-signature (channel-major) mirrors live in ``dh3d_b200.user_ops`` / ``dh3d_b200.tf_ops``.
Outputs are allocated here with torch (the reference: TF's ``allocate_output``).
"""
import ctypes

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, call, check, opt, query, stream_ptr, workspace

f32, i32 = torch.float32, torch.int32
_ci, _cf, _cs = ctypes.c_int, ctypes.c_float, ctypes.c_size_t


def knn_points(xyz, k, keep_workspace=False):
    """xyz [B,N,3] -> (ids [B,N,K] i32, dists [B,N,K] f32).  Reference order and ties, see knn.cu.
    keep_workspace=True also returns the call's workspace tensor: it starts with the cell-sorted copy of the cloud,
    which ``three_nn(..., sorted1=...)`` can walk instead of sorting the same points again."""
    px = check(xyz, f32, "xyz", 3)
    B, N, D = xyz.shape
    if D != 3:
        raise _lib.Dh3dError("knn_points: last dim must be 3")
    ids = torch.empty((B, N, k), dtype=i32, device=xyz.device)
    dists = torch.empty((B, N, k), dtype=f32, device=xyz.device)
    ws, wp, wn = workspace(query("dh3d_knn_workspace_bytes", B, N), xyz.device)
    _lib.stats.tag = "B%d_N%d_K%d" % (B, N, k)
    call("dh3d_knn_bruteforce_pm", px, B, N, int(k), check(ids, i32, "ids"), check(dists, f32, "dists"),
         wp, wn, stream_ptr(xyz.device))
    _lib.stats.tag = None
    if keep_workspace:
        ws._dh3d_sorted_of = (B, N)
        return ids, dists, ws
    return ids, dists


def knn_sort(xyz):
    """xyz [B,N,3] -> the k-NN engine's workspace holding the cell-sorted copy of the cloud (+ chunk boxes): the input of
    ``knn_query_sorted`` and of the ``sorted``-taking forms of ``farthest_point_sample`` / ``three_nn``."""
    px = check(xyz, f32, "xyz", 3)
    B, N, D = xyz.shape
    if D != 3:
        raise _lib.Dh3dError("knn_sort: last dim must be 3")
    ws, wp, wn = workspace(query("dh3d_knn_workspace_bytes", B, N), xyz.device)
    call("dh3d_knn_sort_pm", px, B, N, wp, wn, stream_ptr(xyz.device))
    ws._dh3d_sorted_of = (B, N)
    return ws


def knn_query_sorted(sorted_ws, k):
    """-> (ids [B,N,K] i32, dists [B,N,K] f32) of the cloud ``knn_sort`` sorted; knn_sort + this == knn_points."""
    B, N = sorted_ws._dh3d_sorted_of
    ids = torch.empty((B, N, k), dtype=i32, device=sorted_ws.device)
    dists = torch.empty((B, N, k), dtype=f32, device=sorted_ws.device)
    _lib.stats.tag = "B%d_N%d_K%d" % (B, N, k)
    call("dh3d_knn_query_sorted", ctypes.c_void_p(sorted_ws.data_ptr()), B, N, int(k), check(ids, i32, "ids"),
         check(dists, f32, "dists"), stream_ptr(sorted_ws.device))
    _lib.stats.tag = None
    return ids, dists


def flex_conv(features, theta, bias, neighborhood, xyz, feature_bias=None, scale=None, shift=None,
              act=ACT_NONE):
    """features [B,N,Din], theta [3,Din,Dout], bias [Din,Dout], neighborhood [B,N,K] i32,
    xyz [B,N,3] -> act((flexconv + feature_bias) * scale + shift)  [B,N,Dout]."""
    B, N, Din = features.shape
    K = neighborhood.shape[2]
    Dout = theta.shape[2]
    if tuple(theta.shape) != (3, Din, Dout) or tuple(bias.shape) != (Din, Dout):
        raise _lib.Dh3dError("flex_conv: theta/bias shapes %s %s do not match Din=%d" %
                             (tuple(theta.shape), tuple(bias.shape), Din))
    if tuple(neighborhood.shape[:2]) != (B, N) or tuple(xyz.shape) != (B, N, 3):
        raise _lib.Dh3dError("flex_conv: neighborhood/xyz shape mismatch")
    out = torch.empty((B, N, Dout), dtype=f32, device=features.device)
    ws, wp, wn = workspace(query("dh3d_flex_conv_pm_workspace_bytes", B, N, K, Din, Dout),
                           features.device)
    _lib.stats.tag = "n%d_K%d_Ci%d_Co%d" % (B * N, K, Din, Dout)
    call("dh3d_flex_conv_pm", check(features, f32, "features"), check(theta, f32, "theta"),
         check(bias, f32, "bias"), check(neighborhood, i32, "neighborhood"), check(xyz, f32, "xyz"),
         check(out, f32, "out"), B, N, K, Din, Dout, opt(feature_bias, f32, "feature_bias"),
         opt(scale, f32, "scale"), opt(shift, f32, "shift"), int(act), wp, wn,
         stream_ptr(features.device))
    _lib.stats.tag = None
    return out


def flex_pool(features, neighborhood, with_argmax=False):
    B, N, D = features.shape
    K = neighborhood.shape[2]
    out = torch.empty_like(features)
    arg = torch.empty((B, N, D), dtype=i32, device=features.device) if with_argmax else None
    _lib.stats.tag = "n%d_K%d_D%d_A%d" % (B * N, K, D, int(with_argmax))
    call("dh3d_flex_pool_pm", check(features, f32, "features", 3),
         check(neighborhood, i32, "neighborhood", 3), check(out, f32, "out"), opt(arg, i32, "argmax"),
         B, N, K, D, stream_ptr(features.device))
    _lib.stats.tag = None
    return (out, arg) if with_argmax else out


def flex_conv_prepack(theta, bias, feature_bias=None, scale=None, shift=None):
    """theta [3,Din,Dout], bias [Din,Dout] (+ feature bias / folded BN) -> opaque weight buffer for
    ``flex_conv_packed`` (done once per layer; the buffer remembers its (Din, Dout))."""
    _, Din, Dout = theta.shape
    if tuple(bias.shape) != (Din, Dout):
        raise _lib.Dh3dError("flex_conv_prepack: bias shape %s does not match theta" % (tuple(bias.shape),))
    packed = torch.empty(query("dh3d_flex_conv_prepack_bytes", Din, Dout), dtype=torch.uint8, device=theta.device)
    call("dh3d_flex_conv_prepack", check(theta, f32, "theta", 3), check(bias, f32, "bias", 2),
         opt(feature_bias, f32, "feature_bias"), opt(scale, f32, "scale"), opt(shift, f32, "shift"), Din, Dout,
         ctypes.c_void_p(packed.data_ptr()), stream_ptr(theta.device))
    packed._dh3d_dims = (Din, Dout)
    return packed


def flex_conv_packed(features, packed, neighborhood, xyz, scale=None, act=ACT_NONE):
    """flex_conv with the weights given as flex_conv_prepack(...); ``scale`` must be the one packed."""
    B, N, Din = features.shape
    K = neighborhood.shape[2]
    Dp, Dout = packed._dh3d_dims
    if Dp != Din:
        raise _lib.Dh3dError("flex_conv_packed: features have %d channels, weights expect %d" % (Din, Dp))
    if tuple(neighborhood.shape[:2]) != (B, N) or tuple(xyz.shape) != (B, N, 3):
        raise _lib.Dh3dError("flex_conv_packed: neighborhood/xyz shape mismatch")
    out = torch.empty((B, N, Dout), dtype=f32, device=features.device)
    ws, wp, wn = workspace(query("dh3d_flex_conv_pm_packed_workspace_bytes", B, N, K, Din, Dout), features.device)
    _lib.stats.tag = "n%d_K%d_Ci%d_Co%d" % (B * N, K, Din, Dout)
    call("dh3d_flex_conv_pm_packed", check(features, f32, "features"), ctypes.c_void_p(packed.data_ptr()),
         check(neighborhood, i32, "neighborhood"), check(xyz, f32, "xyz"), check(out, f32, "out"),
         B, N, K, Din, Dout, opt(scale, f32, "scale"), int(act), wp, wn, stream_ptr(features.device))
    _lib.stats.tag = None
    return out


def conv_pointset(features, theta, bias, neighborhood, scale=None, shift=None, act=ACT_NONE):
    B, N, Din = features.shape
    K = neighborhood.shape[2]
    Dout = theta.shape[1]
    if tuple(theta.shape) != (Din, Dout) or tuple(bias.shape) != (Dout,):
        raise _lib.Dh3dError("conv_pointset: theta/bias shape mismatch")
    out = torch.empty((B, N, Dout), dtype=f32, device=features.device)
    _lib.stats.tag = "n%d_K%d_Ci%d_Co%d" % (B * N, K, Din, Dout)
    call("dh3d_conv_pointset_pm", check(features, f32, "features", 3), check(theta, f32, "theta"),
         check(bias, f32, "bias"), check(neighborhood, i32, "neighborhood", 3), check(out, f32, "out"),
         B, N, K, Din, Dout, opt(scale, f32, "scale"), opt(shift, f32, "shift"), int(act),
         stream_ptr(features.device))
    _lib.stats.tag = None
    return out


def farthest_point_sample(npoint, inp, sorted_ws=None):
    """inp [B,N,3] -> [B,npoint] i32 (reference order and ties).  sorted_ws: the k-NN workspace of the same cloud
    (``knn_sort`` / ``knn_points(keep_workspace=True)``), N <= 8192: the box-pruned kernel, same indices."""
    B, N, _ = inp.shape
    out = torch.empty((B, npoint), dtype=i32, device=inp.device)
    _lib.stats.tag = "B%d_N%d_M%d" % (B, N, int(npoint))
    if sorted_ws is not None and N <= 8192:
        if getattr(sorted_ws, "_dh3d_sorted_of", None) != (B, N):
            raise _lib.Dh3dError("farthest_point_sample: sorted_ws is not the k-NN workspace of a [%d,%d,3] cloud" % (B, N))
        call("dh3d_farthest_point_sample_presorted", B, N, int(npoint), ctypes.c_void_p(sorted_ws.data_ptr()),
             check(out, i32, "out"), stream_ptr(inp.device))
        _lib.stats.tag = None
        return out
    call("dh3d_farthest_point_sample", B, N, int(npoint), check(inp, f32, "inp", 3),
         check(out, i32, "out"), stream_ptr(inp.device))
    _lib.stats.tag = None
    return out


def gather_point(inp, idx):
    B, N, _ = inp.shape
    M = idx.shape[1]
    out = torch.empty((B, M, 3), dtype=f32, device=inp.device)
    call("dh3d_gather_point", B, N, M, check(inp, f32, "inp", 3), check(idx, i32, "idx", 2),
         check(out, f32, "out"), stream_ptr(inp.device))
    return out


def group_point(points, idx):
    B, N, C = points.shape
    _, M, S = idx.shape
    out = torch.empty((B, M, S, C), dtype=f32, device=points.device)
    _lib.stats.tag = "B%d_M%d_S%d_C%d" % (B, M, S, C)
    call("dh3d_group_point", B, N, C, M, S, check(points, f32, "points", 3), check(idx, i32, "idx", 3),
         check(out, f32, "out"), stream_ptr(points.device))
    _lib.stats.tag = None
    return out


def group_point_grad(points, idx, grad_out):
    """tf_grouping.py:57-61 (_group_point_grad): grad_out [B,M,S,C], idx [B,M,S] -> grad_points [B,N,C]."""
    B, N, C = points.shape
    _, M, S = idx.shape
    gp = torch.empty((B, N, C), dtype=f32, device=grad_out.device)
    call("dh3d_group_point_grad", B, N, C, M, S, check(grad_out, f32, "grad_out", 4), check(idx, i32, "idx", 3),
         check(gp, f32, "grad_points"), stream_ptr(grad_out.device))
    return gp


def gather_point_grad(inp, idx, out_g):
    """tf_sampling.py:47-51 (_gather_point_grad): out_g [B,M,3], idx [B,M] -> inp_g [B,N,3]."""
    B, N, _ = inp.shape
    M = idx.shape[1]
    gp = torch.empty((B, N, 3), dtype=f32, device=out_g.device)
    call("dh3d_gather_point_grad", B, N, M, check(out_g, f32, "out_g", 3), check(idx, i32, "idx", 2),
         check(gp, f32, "inp_g"), stream_ptr(out_g.device))
    return gp


def three_interpolate_grad(points, idx, weight, grad_out):
    """tf_interpolate.py:29-34 (_three_interpolate_grad): grad_out [B,n,c] -> grad_points [B,m,c]."""
    B, m, c = points.shape
    n = idx.shape[1]
    gp = torch.empty((B, m, c), dtype=f32, device=grad_out.device)
    call("dh3d_three_interpolate_grad", B, n, c, m, check(grad_out, f32, "grad_out", 3),
         check(idx, i32, "idx", 3), check(weight, f32, "weight", 3), check(gp, f32, "grad_points"),
         stream_ptr(grad_out.device))
    return gp


def query_ball_point(radius, nsample, xyz1, xyz2):
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = torch.empty((B, m, nsample), dtype=i32, device=xyz1.device)
    cnt = torch.empty((B, m), dtype=i32, device=xyz1.device)
    ws, wp, wn = workspace(query("dh3d_query_ball_point_workspace_bytes", B, m), xyz1.device)
    call("dh3d_query_ball_point", B, n, m, _cf(radius), int(nsample), check(xyz1, f32, "xyz1", 3),
         check(xyz2, f32, "xyz2", 3), check(idx, i32, "idx"), check(cnt, i32, "cnt"), wp, wn,
         stream_ptr(xyz1.device))
    return idx, cnt


def three_nn(xyz1, xyz2, exhaustive=False, sorted1=None, sorted2=None):
    """3 nearest xyz2 points of every xyz1 point (squared distances, reference arithmetic and ties).
    Default: Morton-sorted box-pruned scan (same results); exhaustive=True: the plain tiled scan.
    sorted1 / sorted2: the workspaces ``knn_sort`` / ``knn_points(..., keep_workspace=True)`` returned for xyz1 / xyz2
    (skips sorting those points again; sorted2 needs sorted1)."""
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = torch.empty((B, n, 3), dtype=f32, device=xyz1.device)
    idx = torch.empty((B, n, 3), dtype=i32, device=xyz1.device)
    if exhaustive:
        call("dh3d_three_nn", B, n, m, check(xyz1, f32, "xyz1", 3), check(xyz2, f32, "xyz2", 3),
             check(dist, f32, "dist"), check(idx, i32, "idx"), stream_ptr(xyz1.device))
        return dist, idx
    _lib.stats.tag = "B%d_n%d_m%d" % (B, n, m)
    if sorted1 is not None and getattr(sorted1, "_dh3d_sorted_of", None) != (B, n):
        raise _lib.Dh3dError("three_nn: sorted1 is not the k-NN workspace of a [%d,%d,3] cloud" % (B, n))
    if sorted2 is not None and (sorted1 is None or getattr(sorted2, "_dh3d_sorted_of", None) != (B, m)):
        raise _lib.Dh3dError("three_nn: sorted2 must be the k-NN workspace of a [%d,%d,3] cloud, next to sorted1" % (B, m))
    if sorted2 is not None:
        call("dh3d_three_nn_presorted2", B, n, m, ctypes.c_void_p(sorted1.data_ptr()),
             ctypes.c_void_p(sorted2.data_ptr()), check(dist, f32, "dist"), check(idx, i32, "idx"),
             stream_ptr(xyz1.device))
        _lib.stats.tag = None
        return dist, idx
    ws, wp, wn = workspace(query("dh3d_three_nn_workspace_bytes", B, n, m), xyz1.device)
    if sorted1 is not None:
        call("dh3d_three_nn_ws_presorted", B, n, m, ctypes.c_void_p(sorted1.data_ptr()), check(xyz2, f32, "xyz2", 3),
             check(dist, f32, "dist"), check(idx, i32, "idx"), wp, wn, stream_ptr(xyz1.device))
        _lib.stats.tag = None
        return dist, idx
    call("dh3d_three_nn_ws", B, n, m, check(xyz1, f32, "xyz1", 3), check(xyz2, f32, "xyz2", 3),
         check(dist, f32, "dist"), check(idx, i32, "idx"), wp, wn, stream_ptr(xyz1.device))
    _lib.stats.tag = None
    return dist, idx


def three_interpolate(points, idx, weight, weight_is_dist2=False, out=None, out_col=0):
    """``out``/``out_col``: write into columns [out_col, out_col+c) of an existing [B,n,ld] tensor (fused concat)."""
    B, m, c = points.shape
    n = idx.shape[1]
    _lib.stats.tag = "B%d_n%d_m%d_C%d" % (B, n, m, c)
    if out is not None:
        check(out, f32, "out", 3)
        call("dh3d_three_interpolate_ld", B, m, c, n, check(points, f32, "points", 3), check(idx, i32, "idx", 3),
             check(weight, f32, "weight", 3), int(bool(weight_is_dist2)),
             ctypes.c_void_p(out.data_ptr() + 4 * out_col), out.shape[-1], stream_ptr(points.device))
        _lib.stats.tag = None
        return out
    out = torch.empty((B, n, c), dtype=f32, device=points.device)
    name = "dh3d_three_interpolate_from_dist" if weight_is_dist2 else "dh3d_three_interpolate"
    call(name, B, m, c, n, check(points, f32, "points", 3), check(idx, i32, "idx", 3),
         check(weight, f32, "weight", 3), check(out, f32, "out"), stream_ptr(points.device))
    _lib.stats.tag = None
    return out


def group_point_cols(wide, col, c, idx):
    """group_point on the column block [col, col+c) of ``wide`` [B,N,ld]: -> [B,M,S,c]."""
    B, N, ld = wide.shape
    _, M, S = idx.shape
    out = torch.empty((B, M, S, c), dtype=f32, device=wide.device)
    check(wide, f32, "wide", 3)
    _lib.stats.tag = "B%d_M%d_S%d_C%d" % (B, M, S, c)
    call("dh3d_group_point_ld", B, N, c, M, S, ctypes.c_void_p(wide.data_ptr() + 4 * col), ld,
         check(idx, i32, "idx", 3), check(out, f32, "out"), stream_ptr(wide.device))
    _lib.stats.tag = None
    return out


def add_l2_normalize_rows(a, b, eps):
    """(a + b, l2-normalised rows of a + b) in one pass."""
    M, C = _rows(a)
    s, y = torch.empty_like(a), torch.empty_like(a)
    call("dh3d_add_l2_normalize_rows", check(a, f32, "a"), check(b, f32, "b"), check(s, f32, "sum"),
         check(y, f32, "y"), M, C, _cf(float(eps)), stream_ptr(a.device))
    return s, y


def linear_join(xa, packed_a, scale_a, shift_a, act_a, xb, packed_b, scale_b, shift_b, act_b, eps=None, out_norm=None):
    """act_a((xa @ Wa)*scale_a + shift_a) + act_b((xb @ Wb)*scale_b + shift_b), and with ``eps`` also its
    l2-normalised rows, in one launch (weights as linear_prepack buffers; N must be 128).
    Returns y, or (y, y_normalised) when eps is given.  ``out_norm``: caller-owned [.., N] tensor for y_normalised."""
    M, Ka = _rows(xa)
    Mb, Kb = _rows(xb)
    (Kpa, N), (Kpb, Nb) = packed_a._dh3d_kn, packed_b._dh3d_kn
    if M != Mb or Kpa != Ka or Kpb != Kb or N != Nb:
        raise _lib.Dh3dError("linear_join: shape mismatch")
    y = torch.empty(xa.shape[:-1] + (N,), dtype=f32, device=xa.device)
    yn = None
    if eps is not None:
        yn = torch.empty_like(y) if out_norm is None else out_norm
        if tuple(yn.shape) != tuple(y.shape):
            raise _lib.Dh3dError("linear_join: out_norm has shape %s, expected %s" % (tuple(yn.shape), tuple(y.shape)))
    _lib.stats.tag = "M%d_Ka%d_Kb%d_N%d" % (M, Ka, Kb, N)
    call("dh3d_linear_join_packed", check(xa, f32, "xa"), Ka, ctypes.c_void_p(packed_a.data_ptr()),
         opt(scale_a, f32, "scale_a"), opt(shift_a, f32, "shift_a"), int(act_a), check(xb, f32, "xb"), Kb,
         ctypes.c_void_p(packed_b.data_ptr()), opt(scale_b, f32, "scale_b"), opt(shift_b, f32, "shift_b"),
         int(act_b), check(y, f32, "y"), N, opt(yn, f32, "yn"), N, _cf(float(eps or 0.0)), M, Ka, Kb, N,
         stream_ptr(xa.device))
    _lib.stats.tag = None
    return (y, yn) if eps is not None else y


def linear_chain_supported(K1, N1, N2):
    """Shapes ``linear_chain`` is built for (dh3d_linear_chain_packed): K1 <= 128, N1 == 128, N2 <= 256."""
    return 0 < K1 <= 128 and K1 % 4 == 0 and N1 == 128 and 0 < N2 <= 256 and N2 % 4 == 0


def linear_chain(x, packed1, scale1, shift1, act1, packed2, scale2, shift2, act2):
    """act2((act1((x @ W1)*scale1 + shift1) @ W2)*scale2 + shift2) in one launch; the hidden activation never
    reaches HBM.  Weights as linear_prepack buffers of W1 [K1,128], W2 [128,N2]."""
    M, K1 = _rows(x)
    (Kp1, N1), (Kp2, N2) = packed1._dh3d_kn, packed2._dh3d_kn
    if Kp1 != K1 or Kp2 != N1 or not linear_chain_supported(K1, N1, N2):
        raise _lib.Dh3dError("linear_chain: unsupported shapes K1=%d N1=%d (W2 rows %d) N2=%d" % (K1, N1, Kp2, N2))
    y = torch.empty(x.shape[:-1] + (N2,), dtype=f32, device=x.device)
    _lib.stats.tag = "M%d_K%d_H%d_N%d" % (M, K1, N1, N2)
    call("dh3d_linear_chain_packed", check(x, f32, "x"), K1, ctypes.c_void_p(packed1.data_ptr()),
         opt(scale1, f32, "scale1"), opt(shift1, f32, "shift1"), int(act1), ctypes.c_void_p(packed2.data_ptr()),
         opt(scale2, f32, "scale2"), opt(shift2, f32, "shift2"), int(act2), check(y, f32, "y"), N2, M, K1, N1, N2,
         stream_ptr(x.device))
    _lib.stats.tag = None
    return y


def _rows(x):
    """[..., C] contiguous tensor -> (M, C) row view parameters."""
    C = x.shape[-1]
    return x.numel() // C, C


def linear_prepack(w):
    """w [K,N] -> opaque packed buffer {W_hi^T, W_lo^T} for the tensor-core path (done once)."""
    K, N = w.shape
    packed = torch.empty(query("dh3d_linear_prepack_bytes", K, N), dtype=torch.uint8, device=w.device)
    call("dh3d_linear_prepack", check(w, f32, "w", 2), K, N, ctypes.c_void_p(packed.data_ptr()),
         stream_ptr(w.device))
    packed._dh3d_kn = (K, N)
    return packed


def linear(x, w, scale=None, shift=None, act=ACT_NONE, out=None, out_col=0, packed=None):
    """act((x @ w) * scale + shift) over the last dim of x.  ``out``/``out_col`` write the result
    into columns [out_col, out_col+N) of an existing [..., ldy] tensor (fused concat).
    ``packed`` (from linear_prepack(w)) selects the tcgen05 3xTF32 kernel; otherwise fp32 FFMA."""
    M, K = _rows(x)
    N = w.shape[1]
    if w.shape[0] != K:
        raise _lib.Dh3dError("linear: x has %d columns, w has %d rows" % (K, w.shape[0]))
    if packed is not None and getattr(packed, "_dh3d_kn", (K, N)) != (K, N):
        raise _lib.Dh3dError("linear: packed weight does not match w")
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), dtype=f32, device=x.device)
        ldy, yptr = N, check(out, f32, "out")
    else:
        check(out, f32, "out")
        ldy = out.shape[-1]
        yptr = ctypes.c_void_p(out.data_ptr() + 4 * out_col)
    _lib.stats.tag = "M%d_K%d_N%d" % (M, K, N)
    if packed is not None:
        call("dh3d_linear_packed", check(x, f32, "x"), K, ctypes.c_void_p(packed.data_ptr()),
             opt(scale, f32, "scale"), opt(shift, f32, "shift"), int(act), yptr, ldy, M, K, N,
             stream_ptr(x.device))
    else:
        call("dh3d_linear", check(x, f32, "x"), K, check(w, f32, "w", 2), opt(scale, f32, "scale"),
             opt(shift, f32, "shift"), int(act), yptr, ldy, M, K, N, stream_ptr(x.device))
    _lib.stats.tag = None
    return out


def linear_rowdot(x, packed, scale, shift, act, w2, b2=0.0, act2=ACT_NONE, out=None):
    """act2(sum_n act((x @ W)[.., n]*scale+shift) * w2[n] + b2) with W given as linear_prepack(W);
    the hidden [.., N] activation is never written (tensor-core path only)."""
    M, K = _rows(x)
    Kp, N = packed._dh3d_kn
    if Kp != K or w2.numel() != N:
        raise _lib.Dh3dError("linear_rowdot: shape mismatch")
    y = torch.empty(x.shape[:-1], dtype=f32, device=x.device) if out is None else out
    if y.numel() != M:
        raise _lib.Dh3dError("linear_rowdot: out has %d elements, expected %d" % (y.numel(), M))
    _lib.stats.tag = "M%d_K%d_N%d" % (M, K, N)
    call("dh3d_linear_rowdot_packed", check(x, f32, "x"), K, ctypes.c_void_p(packed.data_ptr()),
         opt(scale, f32, "scale"), opt(shift, f32, "shift"), int(act), check(w2, f32, "w2"),
         _cf(float(b2)), int(act2), check(y, f32, "y"), M, K, N, stream_ptr(x.device))
    _lib.stats.tag = None
    return y


def rowdot(x, w, bias=0.0, act=ACT_NONE):
    M, K = _rows(x)
    y = torch.empty(x.shape[:-1], dtype=f32, device=x.device)
    call("dh3d_rowdot", check(x, f32, "x"), K, check(w, f32, "w"), _cf(float(bias)), int(act),
         check(y, f32, "y"), M, K, stream_ptr(x.device))
    return y


def se_excite(x, gate):
    y = torch.empty_like(x)
    call("dh3d_se_excite", check(x, f32, "x"), check(gate, f32, "gate"), check(y, f32, "y"),
         _cs(x.numel()), stream_ptr(x.device))
    return y


def se_pool_excite(x, neighborhood, w1, b1, w2, b2):
    """relu(x + x * sigmoid(relu(flex_pool(x, nbr) @ w1 + b1) @ w2 + b2)) in one launch
    (se_res_bottleneck, core/backbones.py:45-55).  x [B,N,C], neighborhood [B,N,K], C in {64,128}, H = C/4."""
    B, N, C = x.shape
    K = neighborhood.shape[2]
    H = w1.shape[1]
    y = torch.empty_like(x)
    _lib.stats.tag = "n%d_K%d_C%d" % (B * N, K, C)
    call("dh3d_se_pool_excite", check(x, f32, "x", 3), check(neighborhood, i32, "neighborhood", 3),
         check(w1, f32, "w1", 2), check(b1, f32, "b1"), check(w2, f32, "w2", 2), check(b2, f32, "b2"),
         check(y, f32, "y"), B, N, K, C, H, stream_ptr(x.device))
    _lib.stats.tag = None
    return y


def affine(x, a, b):
    """a * x + b elementwise (the reference's `1 - attention`, localdesc_extract.py:95)."""
    y = torch.empty_like(x)
    call("dh3d_affine", check(x, f32, "x"), _cf(float(a)), _cf(float(b)), check(y, f32, "y"), _cs(x.numel()),
         stream_ptr(x.device))
    return y


def gather_rows_cols(src, idx, out, out_col):
    """out[b, j, out_col:out_col+C] = src[b, idx[b,j], :] (zero where idx < 0); src [B,N,C], idx [B,M] i32,
    out [B,M,ld] (the reference's `res[max_indices, :]`, written per column block)."""
    B, N, C = src.shape
    M = idx.shape[1]
    check(out, f32, "out", 3)
    if out.shape[0] != B or out.shape[1] != M or out_col + C > out.shape[2]:
        raise _lib.Dh3dError("gather_rows_cols: out shape %s does not fit" % (tuple(out.shape),))
    call("dh3d_gather_rows", B, N, C, M, check(src, f32, "src", 3), check(idx, i32, "idx", 2),
         ctypes.c_void_p(out.data_ptr() + 4 * out_col), out.shape[2], stream_ptr(src.device))
    return out


def add(a, b):
    y = torch.empty_like(a)
    call("dh3d_add", check(a, f32, "a"), check(b, f32, "b"), check(y, f32, "y"), _cs(a.numel()),
         stream_ptr(a.device))
    return y


def l2_normalize_rows(x, eps, out=None, out_col=0):
    M, C = _rows(x)
    if out is None:
        out = torch.empty_like(x)
        ldy, yptr = C, check(out, f32, "out")
    else:
        check(out, f32, "out")
        ldy = out.shape[-1]
        yptr = ctypes.c_void_p(out.data_ptr() + 4 * out_col)
    call("dh3d_l2_normalize_rows", check(x, f32, "x"), C, yptr, ldy, M, C, _cf(eps),
         stream_ptr(x.device))
    return out


def copy_cols(src, dst, dst_col):
    """dst[..., dst_col:dst_col+C] = src[..., :C]"""
    M, C = _rows(src)
    check(dst, f32, "dst")
    call("dh3d_copy_cols", check(src, f32, "src"), C, ctypes.c_void_p(dst.data_ptr() + 4 * dst_col),
         dst.shape[-1], M, C, stream_ptr(src.device))
    return dst


def transpose_cm_to_pm(x):
    """[B,C,N] -> [B,N,C] (fp32 or int32)."""
    B, C, N = x.shape
    out = torch.empty((B, N, C), dtype=x.dtype, device=x.device)
    call("dh3d_transpose_cm_to_pm", check(x, x.dtype, "x", 3), check(out, x.dtype, "out"), B, C, N,
         stream_ptr(x.device))
    return out


def transpose_pm_to_cm(x):
    B, N, C = x.shape
    out = torch.empty((B, C, N), dtype=x.dtype, device=x.device)
    call("dh3d_transpose_pm_to_cm", check(x, x.dtype, "x", 3), check(out, x.dtype, "out"), B, N, C,
         stream_ptr(x.device))
    return out


def netvlad(features, att, cluster_weights, cluster_bn, cluster_weights2, hidden1_weights, bn,
            gating_weights, gating_bn, final_l2norm=True, out=None):
    """features [B,N,D], att [B,N] (or [B,N,1]); *_bn = (scale, shift) folded BatchNorm pairs."""
    B, N, D = features.shape
    Kc = cluster_weights.shape[1]
    out_dim = hidden1_weights.shape[1]
    att = att.reshape(B, N)
    if out is None:
        out = torch.empty((B, out_dim), dtype=f32, device=features.device)
    elif tuple(out.shape) != (B, out_dim):
        raise _lib.Dh3dError("netvlad: out has shape %s, expected %s" % (tuple(out.shape), (B, out_dim)))
    nbytes = query("dh3d_netvlad_workspace_bytes", B, N, D, Kc, out_dim)
    if nbytes == 0:
        raise _lib.Dh3dError("netvlad: unsupported configuration D=%d Kc=%d out=%d" % (D, Kc, out_dim))
    ws, wp, wn = workspace(nbytes, features.device)
    _lib.stats.tag = "B%d_N%d_D%d_Kc%d_O%d" % (B, N, D, Kc, out_dim)
    call("dh3d_netvlad", check(features, f32, "features", 3), check(att, f32, "att"), B, N, D, Kc,
         out_dim, check(cluster_weights, f32, "cluster_weights"), check(cluster_bn[0], f32, "cbn_s"),
         check(cluster_bn[1], f32, "cbn_b"), check(cluster_weights2.reshape(D, Kc), f32, "cw2"),
         check(hidden1_weights, f32, "hidden1_weights"), check(bn[0], f32, "bn_s"),
         check(bn[1], f32, "bn_b"), check(gating_weights, f32, "gating_weights"),
         check(gating_bn[0], f32, "gbn_s"), check(gating_bn[1], f32, "gbn_b"), int(bool(final_l2norm)),
         check(out, f32, "out"), wp, wn, stream_ptr(features.device))
    _lib.stats.tag = None
    return out
