"""Layer-level API (boundary C): the parameter-owning layers ``core/model.py`` builds on.

Mirrors ``core/layers.py`` of the reference (Keras ``Layer`` classes + function forms) as
``torch.nn.Module``s whose parameters carry the reference's variable names and shapes, so a
name-mapped loader can fill them from the shipped TF checkpoints:

    KnnBruteforce / knn_bruteforce            core/layers.py:49-107   (returns [B,K,N] like the layer)
    FlexPooling / flex_pooling                :110-175
    FlexConvolution / flex_convolution        :178-339, 439-461   position_theta [3,Din,Dout],
                                              position_bias [Din,Dout], feature_bias [Dout,1]
    Flex_Avg / flex_avg                       :342-436, 464-480   theta = 0, bias = I
    ConvolutionPointset / convolution_pointset :564-707           position_theta [Din,Dout], position_bias [Dout]

``forward`` takes the reference's channel-major tensors ([B,C,N], [B,K,N]); ``forward_pm`` is the
native point-major fast path the assembled forward pass uses (with the follow-up BatchNorm and
activation fused into the kernel epilogue).  Inference only.
"""
import os

import torch
from torch import nn

from . import ops, user_ops
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID

def use_tensor_cores():
    """DH3D_EXACT_FP32=1 -- the library's one process-wide switch (csrc/capi.cu) -- selects the exact-fp32 debug
    path (FFMA GEMMs, two-kernel FlexConv, FFMA NetVLAD); default is the tcgen05 kernels."""
    return os.environ.get("DH3D_EXACT_FP32", "0") in ("", "0")


TENSORPACK_BN_EPS = 1e-5  # tensorpack BatchNorm default epsilon (library default; parity unpinned)
SLIM_BN_EPS = 1e-3        # tf.contrib slim / layers batch_norm default epsilon


class FoldedModule(nn.Module):
    """Base of every layer that caches weight-derived operands in ``self._folded`` (folded BatchNorm, prepacked
    tensor-core weights, scalars read back to the host).  The cache is keyed on nothing but the parameters'
    current values and device, so everything that can change either drops it: ``.to()/.cuda()/.float()``
    (``_apply``), ``load_state_dict`` (``_load_from_state_dict``) and ``invalidate_folded`` for in-place edits."""

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._folded = None
        return out

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._folded = None


class BatchNorm(nn.Module):
    """Inference BatchNorm statistics.  tensorpack names: gamma, beta, mean/EMA, variance/EMA;
    slim names: gamma, beta, moving_mean, moving_variance -- both map to these four tensors."""

    def __init__(self, channels, eps=TENSORPACK_BN_EPS):
        super().__init__()
        self.eps = eps
        self.gamma = nn.Parameter(torch.ones(channels), requires_grad=False)
        self.beta = nn.Parameter(torch.zeros(channels), requires_grad=False)
        self.mean_ema = nn.Parameter(torch.zeros(channels), requires_grad=False)
        self.variance_ema = nn.Parameter(torch.ones(channels), requires_grad=False)

    def fold(self, pre_bias=None):
        """(scale, shift) with y = x*scale + shift == BN(x + pre_bias)."""
        scale = self.gamma / torch.sqrt(self.variance_ema + self.eps)
        bias = 0.0 if pre_bias is None else pre_bias.reshape(-1)
        shift = (bias - self.mean_ema) * scale + self.beta
        return scale.contiguous(), shift.contiguous()


def invalidate_folded(module):
    """Drop every cached weight-derived operand (call after editing parameters in place; device moves and
    ``load_state_dict`` do it themselves, see FoldedModule)."""
    for m in module.modules():
        if hasattr(m, "_folded"):
            m._folded = None
        if hasattr(m, "_side"):
            m._side = None


class VariableStore(nn.ModuleDict):
    """What a TF variable scope is to the reference's function forms (``layer.apply`` under
    ``tf.variable_scope``): the layer a function form creates is kept under its ``name`` and re-used by later
    calls with the same name, so its parameters can be filled by a checkpoint loader and moved with ``.cuda()``."""

    def layer(self, name, factory):
        if name is None:
            raise ValueError("the function forms need name= (the reference's variable scope) to own parameters")
        key = name.replace(".", "/")
        if key not in self:
            self[key] = factory()
        return self[key]


default_store = VariableStore()


def _activation(y, activation):
    """Keras ``activations.get``: None / 'linear' / callable / 'relu' / 'sigmoid'."""
    if activation is None or activation == "linear":
        return y
    if callable(activation):
        return activation(y)
    if activation == "relu":
        return torch.relu(y)
    if activation == "sigmoid":
        return torch.sigmoid(y)
    raise ValueError("unsupported activation %r" % (activation,))


class KnnBruteforce(nn.Module):
    def __init__(self, k, data_format="simple"):
        super().__init__()
        assert k > 0 and data_format == "simple"
        self.k = k

    def forward(self, positions):
        """positions [B,Dp,N] -> (NN [B,K,N] i32, distances [B,K,N])  (core/layers.py:85-98)."""
        nn_, dist = user_ops.knn_bruteforce(positions, self.k)
        return ops.transpose_pm_to_cm(nn_), ops.transpose_pm_to_cm(dist)


def knn_bruteforce(positions, k, data_format="simple", name=None):
    return KnnBruteforce(k, data_format)(positions)


class FlexPooling(nn.Module):
    def forward(self, features, neighborhoods):
        return user_ops.flex_pooling(features, neighborhoods)[0]

    def forward_pm(self, features, neighborhoods):
        return ops.flex_pool(features, neighborhoods)


def flex_pooling(features, neighborhoods, data_format="simple", name=None):
    return FlexPooling()(features, neighborhoods)


class FlexConvolution(FoldedModule):
    def __init__(self, in_channels, filters, use_feature_bias=True, dp=3):
        super().__init__()
        self.filters = int(filters)
        self.position_theta = nn.Parameter(torch.zeros(dp, in_channels, filters), requires_grad=False)
        self.position_bias = nn.Parameter(torch.zeros(in_channels, filters), requires_grad=False)
        self.feature_bias = (nn.Parameter(torch.zeros(filters, 1), requires_grad=False)
                             if use_feature_bias else None)
        self._folded = None

    def forward(self, features, positions, neighborhoods):
        y = user_ops.flex_convolution(features, positions, neighborhoods, self.position_theta,
                                      self.position_bias)
        if self.feature_bias is not None:
            y = y + self.feature_bias  # [Dout,1] broadcast over [B,Dout,N]  (core/layers.py:330-331)
        return y

    def forward_pm(self, features, xyz, neighborhoods, bn=None, act=ACT_NONE):
        if self._folded is None:  # folded once; call invalidate_folded(model) after loading weights
            scale = shift = None
            if bn is not None:
                scale, shift = bn.fold()
            fb = None if self.feature_bias is None else self.feature_bias.reshape(-1).contiguous()
            # weight-only operand prepared once (dh3d_flex_conv_prepack): no per-forward packing launches
            packed = (ops.flex_conv_prepack(self.position_theta, self.position_bias, fb, scale, shift)
                      if self.position_theta.is_cuda else None)
            self._folded = (fb, scale, shift, packed)
        fb, scale, shift, packed = self._folded
        if packed is not None:
            return ops.flex_conv_packed(features, packed, neighborhoods, xyz, scale=scale, act=act)
        return ops.flex_conv(features, self.position_theta, self.position_bias, neighborhoods, xyz,
                             feature_bias=fb, scale=scale, shift=shift, act=act)


def flex_convolution(features, positions, neighborhoods, filters, activation=None, kernel_initializer=None,
                     position_bias_initializer=None, features_bias_initializer=None, use_feature_bias=True,
                     data_format="simple", trainable=True, name=None, store=None):
    """core/layers.py:439-461, same argument order: features [B,Din,N], positions [B,Dp,N], neighborhoods
    [B,K,N] i32 -> activation(flexconv + feature_bias) [B,filters,N].  The layer (position_theta [Dp,Din,filters],
    position_bias [Din,filters], feature_bias [filters,1]) lives in ``store`` under ``name``; a new layer starts
    with Glorot-uniform theta and zero biases like the reference's default initialisers (``kernel_initializer``
    may be a callable ``f(tensor)`` applied in place instead)."""
    assert data_format == "simple", "only data_format='simple' (rank-3 tensors) is built"

    def make():
        Din, Dp = features.shape[1], positions.shape[1]
        layer = FlexConvolution(Din, filters, use_feature_bias=use_feature_bias, dp=Dp)
        with torch.no_grad():
            if callable(kernel_initializer):
                kernel_initializer(layer.position_theta)
            else:
                nn.init.xavier_uniform_(layer.position_theta)
        return layer.to(features.device)

    layer = (store if store is not None else default_store).layer(name, make)
    return _activation(layer(features, positions, neighborhoods), activation)


class Flex_Avg(FlexConvolution):
    """core/layers.py:342-436: FlexConv with position_theta = 0 (a non-trainable VARIABLE, so it is in the
    checkpoint) and position_bias = eye(Dout) (a constant, not a variable): the plain neighbour SUM; the
    caller scales by 1/K (core/backbones.py:80-83).  Requires Din == Dout like the reference's tf.eye(Dout)."""

    def __init__(self, channels, dp=3):
        super().__init__(channels, channels, use_feature_bias=False, dp=dp)
        del self.position_bias
        self.register_buffer("position_bias", torch.eye(channels), persistent=False)


def flex_avg(features, positions, neighborhoods, filters, activation=None, kernel_initializer=None,
             data_format="simple", trainable=True, name=None, store=None):
    """core/layers.py:464-480, same argument order -> neighbour sum [B,filters,N] (filters must equal Din)."""
    assert data_format == "simple"
    if int(filters) != features.shape[1]:
        raise ValueError("flex_avg: filters (%d) must equal the input channels (%d): position_bias is eye(filters)"
                         % (filters, features.shape[1]))
    layer = (store if store is not None else default_store).layer(
        name, lambda: Flex_Avg(int(filters), dp=positions.shape[1]).to(features.device))
    return _activation(layer(features, positions, neighborhoods), activation)


class ConvolutionPointset(FoldedModule):
    def __init__(self, in_channels, filters, use_feature_bias=False):
        super().__init__()
        self.filters = int(filters)
        self.position_theta = nn.Parameter(torch.zeros(in_channels, filters), requires_grad=False)
        self.position_bias = nn.Parameter(torch.zeros(filters), requires_grad=False)
        self.feature_bias = (nn.Parameter(torch.zeros(filters, 1), requires_grad=False)
                             if use_feature_bias else None)
        self._folded = None

    def forward(self, features, neighborhoods):
        y = user_ops.convolution_pointset(features, neighborhoods, self.position_theta,
                                          self.position_bias)
        if self.feature_bias is not None:
            y = y + self.feature_bias
        return y

    def forward_pm(self, features, neighborhoods, bn=None, act=ACT_NONE):
        if self._folded is None:
            scale = shift = None
            if bn is not None:
                scale, shift = bn.fold(None if self.feature_bias is None else self.feature_bias)
            self._folded = (scale, shift)
        scale, shift = self._folded
        return ops.conv_pointset(features, self.position_theta, self.position_bias, neighborhoods,
                                 scale=scale, shift=shift, act=act)


def convolution_pointset(features, neighborhoods, filters, activation=None, kernel_initializer=None,
                         position_bias_initializer=None, features_bias_initializer=None, use_feature_bias=False,
                         data_format="simple", trainable=True, name=None, store=None):
    """core/layers.py:686-707, same argument order: features [B,Din,N], neighborhoods [B,K,N] ->
    activation(conv_pointset (+ feature_bias)) [B,filters,N]; parameters owned by ``store[name]``."""
    assert data_format == "simple"

    def make():
        layer = ConvolutionPointset(features.shape[1], filters, use_feature_bias=use_feature_bias)
        with torch.no_grad():
            if callable(kernel_initializer):
                kernel_initializer(layer.position_theta)
            else:
                nn.init.xavier_uniform_(layer.position_theta)
        return layer.to(features.device)

    layer = (store if store is not None else default_store).layer(name, make)
    return _activation(layer(features, neighborhoods), activation)


class Conv1x1(FoldedModule):
    """tensorpack ``Conv2D(kernel_shape=1)`` on [B,N,1,C] (core/tf_utils.py:99-109): W [1,1,Cin,Cout],
    b [Cout], optional BatchNorm ('bn' scope) and activation.  Point-major only."""

    def __init__(self, in_channels, out_channels, bn=True, act=ACT_RELU):
        super().__init__()
        self.W = nn.Parameter(torch.zeros(1, 1, in_channels, out_channels), requires_grad=False)
        self.b = nn.Parameter(torch.zeros(out_channels), requires_grad=False)
        self.bn = BatchNorm(out_channels) if bn else None
        self.act = act
        self._folded = None

    def folded(self):
        if self._folded is None:
            w = self.W.reshape(self.W.shape[2], self.W.shape[3]).contiguous()
            if self.bn is not None:
                scale, shift = self.bn.fold(self.b)
            else:
                scale, shift = None, self.b.detach().contiguous()
            packed = ops.linear_prepack(w) if (use_tensor_cores() and w.is_cuda) else None
            self._folded = (w, scale, shift, packed)
        return self._folded

    def forward(self, x, out=None, out_col=0):
        w, scale, shift, packed = self.folded()
        return ops.linear(x, w, scale=scale, shift=shift, act=self.act, out=out, out_col=out_col,
                          packed=packed)
