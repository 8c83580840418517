"""Drop-in mirror of the reference's ``user_ops`` python functions (boundary A).

Same names, argument order, layouts and return tuples as ``user_ops/__init__.py`` of the
reference, over CUDA ``torch.Tensor`` instead of TF tensors:

    knn_bruteforce(positions[B,Dp,N], k) -> (neighborhood_out[B,N,K] i32, distances[B,N,K])   :50
    flex_convolution(features[B,Din,N], position[B,Dp,N], neighborhood[B,K,N], theta[Dp,Din,Dout],
                     bias[Din,Dout]) -> [B,Dout,N]                                           :63-89
    flex_pooling(features[B,D,N], neighborhood[B,K,N]) -> (max[B,D,N], argmax[B,D,N] i32)     :115-135
    convolution_pointset(features[B,Din,N], neighborhood[B,K,N], theta[Din,Dout], bias[Dout])
                     -> [B,Dout,N]                                                           :205-225

    flex_convolution_transpose(features, position, neighborhood, theta, bias) -> [B,Dout,N]      :155-179

Shape errors raise (the reference: TF shape functions at graph-build time, e.g.
user_ops/ops/flex_conv.cc:42-60).

Gradients: the reference registers ``FlexConvGrad`` / ``FlexPoolGrad`` / ``ConvPointsetGrad`` with TF
(:95-111, 141-151, 231-246).  Here the same three kernels are exposed as ``flex_convolution_grad``,
``flex_pooling_grad``, ``convolution_pointset_grad`` (op-input order of the reference's ``_*_grad``
ops) and wired into ``torch.autograd``: calling the forward functions on tensors that require grad
records a node whose backward launches them (``[df, dtheta, dbias, None, None]`` like :111).
``FlexDeconvGrad`` (:185-201) is not built (``flex_convolution_transpose`` is forward only; DH3D never
calls it).
"""
import torch

from . import _lib
from ._lib import call, check, query, stream_ptr, workspace

__all__ = ["knn_bruteforce", "flex_convolution", "flex_pooling", "convolution_pointset",
           "conv_relative", "flex_convolution_transpose", "flex_convolution_grad", "flex_pooling_grad",
           "convolution_pointset_grad"]

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


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def flex_convolution(features, position, neighborhood, theta, bias, name=None):
    if _needs_grad(features, theta, bias):
        return _FlexConvFn.apply(features, position, neighborhood, theta, bias)
    return _flex_convolution_fwd(features, position, neighborhood, theta, bias)


def _flex_convolution_fwd(features, position, neighborhood, theta, bias):
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
    if _needs_grad(features):
        return _FlexPoolFn.apply(features, neighborhood)
    return _flex_pooling_fwd(features, neighborhood)


def _flex_pooling_fwd(features, neighborhood):
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
    if _needs_grad(features, theta, bias):
        return _ConvPointsetFn.apply(features, neighborhood, theta, bias)
    return _convolution_pointset_fwd(features, neighborhood, theta, bias)


def _convolution_pointset_fwd(features, neighborhood, theta, bias):
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


def _flex_shapes(features, position, neighborhood, theta, bias, what):
    check(features, f32, "features", 3)
    check(position, f32, "position", 3)
    check(neighborhood, i32, "neighborhood", 3)
    if theta.dim() == 4 and theta.shape[0] == 1:
        theta = theta[0]
    check(theta, f32, "theta", 3)
    check(bias, f32, "bias", 2)
    B, Din, N = features.shape
    K = neighborhood.shape[1]
    Dp, Din_t, Dout = theta.shape
    _require(Dp == 3 and position.shape[1] == 3, what + ": Dp must be 3")
    _require(Din_t == Din and tuple(bias.shape) == (Din, Dout), what + ": theta/bias do not match Din")
    _require(tuple(neighborhood.shape) == (B, K, N) and tuple(position.shape) == (B, 3, N),
             what + ": neighborhood/position batch or point count mismatch")
    return theta, B, N, K, Din, Dout


def flex_convolution_transpose(features, position, neighborhood, theta, bias, name=None):
    """out[b,:,nbr(k,n)] += W(p[nbr(k,n)] - p[nbr(0,n)]) . features[b,:,nbr(0,n)]  (flex_deconv_kernel.cc:25-70)."""
    theta, B, N, K, Din, Dout = _flex_shapes(features, position, neighborhood, theta, bias,
                                             "flex_convolution_transpose")
    out = torch.empty((B, Dout, N), dtype=f32, device=features.device)
    ws, wp, wn = workspace(query("dh3d_flex_deconv_workspace_bytes", B, N, K, Din, Dout), features.device)
    call("dh3d_flex_deconv", check(features, f32, "features"), check(theta, f32, "theta"),
         check(bias, f32, "bias"), check(neighborhood, i32, "neighborhood"), check(position, f32, "position"),
         check(out, f32, "out"), B, N, K, Din, Dout, wp, wn, stream_ptr(features.device))
    return out


def flex_convolution_grad(features, theta, bias, neighborhood, position, topdiff):
    """The reference's ``_flex_conv_grad`` op (input order of user_ops/ops/flex_conv.cc:102-150)
    -> (grad_features [B,Din,N], grad_theta [3,Din,Dout], grad_bias [Din,Dout])."""
    theta, B, N, K, Din, Dout = _flex_shapes(features, position, neighborhood, theta, bias,
                                             "flex_convolution_grad")
    check(topdiff, f32, "topdiff", 3)
    _require(tuple(topdiff.shape) == (B, Dout, N), "flex_convolution_grad: topdiff must be [B,Dout,N]")
    dev = features.device
    gf = torch.empty((B, Din, N), dtype=f32, device=dev)
    gt = torch.empty((3, Din, Dout), dtype=f32, device=dev)
    gb = torch.empty((Din, Dout), dtype=f32, device=dev)
    ws, wp, wn = workspace(query("dh3d_flex_conv_grad_workspace_bytes", B, N, K, Din, Dout), dev)
    call("dh3d_flex_conv_grad", check(features, f32, "features"), check(theta, f32, "theta"),
         check(bias, f32, "bias"), check(neighborhood, i32, "neighborhood"), check(position, f32, "position"),
         check(topdiff, f32, "topdiff"), check(gf, f32, "gf"), check(gt, f32, "gt"), check(gb, f32, "gb"),
         B, N, K, Din, Dout, wp, wn, stream_ptr(dev))
    return gf, gt, gb


def flex_pooling_grad(features, neighborhood, topdiff, argmax):
    """The reference's ``_flex_pool_grad`` op (flex_pool_op.cc; python :141-151) -> grad_features [B,D,N]."""
    check(topdiff, f32, "topdiff", 3)
    check(argmax, i32, "argmax", 3)
    B, D, N = topdiff.shape
    _require(tuple(argmax.shape) == (B, D, N) and tuple(features.shape) == (B, D, N),
             "flex_pooling_grad: features/topdiff/argmax must all be [B,D,N]")
    gf = torch.empty((B, D, N), dtype=f32, device=topdiff.device)
    call("dh3d_flex_pool_grad", check(topdiff, f32, "topdiff"), check(argmax, i32, "argmax"),
         check(gf, f32, "gf"), B, N, D, stream_ptr(topdiff.device))
    return gf


def convolution_pointset_grad(features, theta, bias, neighborhood, topdiff):
    """The reference's ``_conv_pointset_grad`` op -> (grad_features [B,Din,N], grad_theta [Din,Dout],
    grad_bias [Dout])."""
    check(features, f32, "features", 3)
    check(neighborhood, i32, "neighborhood", 3)
    if theta.dim() == 3 and theta.shape[0] == 1:
        theta = theta[0]
    check(theta, f32, "theta", 2)
    check(topdiff, f32, "topdiff", 3)
    B, Din, N = features.shape
    K = neighborhood.shape[1]
    Dout = theta.shape[1]
    _require(theta.shape[0] == Din and tuple(topdiff.shape) == (B, Dout, N) and
             tuple(neighborhood.shape) == (B, K, N), "convolution_pointset_grad: shape mismatch")
    dev = features.device
    gf = torch.empty((B, Din, N), dtype=f32, device=dev)
    gt = torch.empty((Din, Dout), dtype=f32, device=dev)
    gb = torch.empty((Dout,), dtype=f32, device=dev)
    ws, wp, wn = workspace(query("dh3d_conv_pointset_grad_workspace_bytes", B, N, K, Din, Dout), dev)
    call("dh3d_conv_pointset_grad", check(features, f32, "features"), check(theta, f32, "theta"),
         check(neighborhood, i32, "neighborhood"), check(topdiff, f32, "topdiff"), check(gf, f32, "gf"),
         check(gt, f32, "gt"), check(gb, f32, "gb"), B, N, K, Din, Dout, wp, wn, stream_ptr(dev))
    return gf, gt, gb


class _FlexConvFn(torch.autograd.Function):
    """@ops.RegisterGradient("FlexConv") (user_ops/__init__.py:95-111): [df, dt, db, None, None]."""

    @staticmethod
    def forward(ctx, features, position, neighborhood, theta, bias):
        ctx.save_for_backward(features, position, neighborhood, theta, bias)
        return _flex_convolution_fwd(features.detach(), position, neighborhood, theta.detach(), bias.detach())

    @staticmethod
    def backward(ctx, topdiff):
        features, position, neighborhood, theta, bias = ctx.saved_tensors
        th = theta[0] if theta.dim() == 4 else theta
        df, dt, db = flex_convolution_grad(features.detach(), th.detach(), bias.detach(), neighborhood,
                                           position, topdiff.contiguous())
        return df, None, None, dt.reshape(theta.shape), db


class _FlexPoolFn(torch.autograd.Function):
    """@ops.RegisterGradient("FlexPool") (:141-151): [df, None]; argmax carries no gradient."""

    @staticmethod
    def forward(ctx, features, neighborhood):
        out, arg = _flex_pooling_fwd(features.detach(), neighborhood)
        ctx.save_for_backward(features, neighborhood, arg)
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, topdiff, _unused):
        features, neighborhood, arg = ctx.saved_tensors
        return flex_pooling_grad(features.detach(), neighborhood, topdiff.contiguous(), arg), None


class _ConvPointsetFn(torch.autograd.Function):
    """@ops.RegisterGradient("ConvPointset") (:231-246): [df, dt, db, None]."""

    @staticmethod
    def forward(ctx, features, neighborhood, theta, bias):
        ctx.save_for_backward(features, neighborhood, theta, bias)
        return _convolution_pointset_fwd(features.detach(), neighborhood, theta.detach(), bias.detach())

    @staticmethod
    def backward(ctx, topdiff):
        features, neighborhood, theta, bias = ctx.saved_tensors
        th = theta[0] if theta.dim() == 3 else theta
        df, dt, db = convolution_pointset_grad(features.detach(), th.detach(), bias.detach(), neighborhood,
                                               topdiff.contiguous())
        return df, None, dt.reshape(theta.shape), db
