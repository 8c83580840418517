"""Drop-in mirror of the reference's ``user_ops`` python functions (boundary A).

Same names, argument order, layouts and return tuples as ``user_ops/__init__.py`` of the
reference, over CUDA ``torch.Tensor`` instead of TF tensors:

    knn_bruteforce(positions[B,Dp,N], k) -> (neighborhood_out[B,N,K] i32, distances[B,N,K])   :50
    flex_convolution(features[B,Din,N], position[B,Dp,N], neighborhood[B,K,N], theta[Dp,Din,Dout],
                     bias[Din,Dout]) -> [B,Dout,N]                                           :63-89
    flex_pooling(features[B,D,N], neighborhood[B,K,N]) -> (max[B,D,N], argmax[B,D,N] i32)     :115-135
    convolution_pointset(features[B,Din,N], neighborhood[B,K,N], theta[Din,Dout], bias[Dout])
                     -> [B,Dout,N]                                                           :205-225

Shape errors raise (the reference: TF shape functions at graph-build time, e.g.
user_ops/ops/flex_conv.cc:42-60).  Forward only; the gradient registrations (:95-111,141-151,
231-246) are out of scope for this round.
"""
import torch

from . import _lib
from ._lib import call, check, query, stream_ptr, workspace

__all__ = ["knn_bruteforce", "flex_convolution", "flex_pooling", "convolution_pointset",
           "conv_relative"]

f32, i32 = torch.float32, torch.int32


def _require(cond, msg):
    if not cond:
        raise _lib.Dh3dError(msg)


def knn_bruteforce(positions, k, name=None):
    check(positions, f32, "positions", 3)
    B, Dp, N = positions.shape
    _require(Dp == 3, "knn_bruteforce: only Dp == 3 is supported (got %d)" % Dp)
    _require(k > 0, "knn_bruteforce: k must be positive")
    ids = torch.empty((B, N, k), dtype=i32, device=positions.device)
    dists = torch.empty((B, N, k), dtype=f32, device=positions.device)
    ws, wp, wn = workspace(query("dh3d_knn_workspace_bytes", B, N), positions.device)
    call("dh3d_knn_bruteforce", check(positions, f32, "positions"), B, Dp, N, int(k),
         check(ids, i32, "ids"), check(dists, f32, "dists"), wp, wn, stream_ptr(positions.device))
    return ids, dists


def flex_convolution(features, position, neighborhood, theta, bias, name=None):
    check(features, f32, "features", 3)
    check(position, f32, "position", 3)
    check(neighborhood, i32, "neighborhood", 3)
    if theta.dim() == 4 and theta.shape[0] == 1:  # the reference docstring's stale leading 1
        theta = theta[0]
    check(theta, f32, "theta", 3)
    check(bias, f32, "bias", 2)
    B, Din, N = features.shape
    K = neighborhood.shape[1]
    Dp, Din_t, Dout = theta.shape
    _require(Dp == 3 and position.shape[1] == 3, "flex_convolution: Dp must be 3")
    _require(Din_t == Din and tuple(bias.shape) == (Din, Dout),
             "flex_convolution: theta %s / bias %s do not match Din=%d" %
             (tuple(theta.shape), tuple(bias.shape), Din))
    _require(tuple(neighborhood.shape) == (B, K, N) and tuple(position.shape) == (B, 3, N),
             "flex_convolution: neighborhood/position batch or point count mismatch")
    out = torch.empty((B, Dout, N), dtype=f32, device=features.device)
    ws, wp, wn = workspace(query("dh3d_flex_conv_workspace_bytes", B, N, K, Din, Dout),
                           features.device)
    call("dh3d_flex_conv", check(features, f32, "features"), check(theta, f32, "theta"),
         check(bias, f32, "bias"), check(neighborhood, i32, "neighborhood"),
         check(position, f32, "position"), check(out, f32, "out"), B, N, K, Din, Dout, wp, wn,
         stream_ptr(features.device))
    return out


def flex_pooling(features, neighborhood, name=None):
    check(features, f32, "features", 3)
    check(neighborhood, i32, "neighborhood", 3)
    B, D, N = features.shape
    K = neighborhood.shape[1]
    _require(tuple(neighborhood.shape) == (B, K, N), "flex_pooling: neighborhood shape mismatch")
    out = torch.empty_like(features)
    arg = torch.empty((B, D, N), dtype=i32, device=features.device)
    call("dh3d_flex_pool", check(features, f32, "features"), check(neighborhood, i32, "neighborhood"),
         check(out, f32, "out"), check(arg, i32, "argmax"), B, N, K, D, stream_ptr(features.device))
    return out, arg


def convolution_pointset(features, neighborhood, theta, bias, name=None):
    check(features, f32, "features", 3)
    check(neighborhood, i32, "neighborhood", 3)
    if theta.dim() == 3 and theta.shape[0] == 1:
        theta = theta[0]
    check(theta, f32, "theta", 2)
    check(bias, f32, "bias", 1)
    B, Din, N = features.shape
    K = neighborhood.shape[1]
    Dout = theta.shape[1]
    _require(theta.shape[0] == Din and bias.shape[0] == Dout,
             "convolution_pointset: theta/bias shape mismatch")
    _require(tuple(neighborhood.shape) == (B, K, N), "convolution_pointset: neighborhood mismatch")
    out = torch.empty((B, Dout, N), dtype=f32, device=features.device)
    call("dh3d_conv_pointset", check(features, f32, "features"), check(theta, f32, "theta"),
         check(bias, f32, "bias"), check(neighborhood, i32, "neighborhood"), check(out, f32, "out"),
         B, N, K, Din, Dout, stream_ptr(features.device))
    return out


conv_relative = convolution_pointset  # the name used in the reference's prose (README.md:76)
