"""Mirror of the forward-pass helpers of the reference's ``core/tf_utils.py`` (boundary C), same names and
argument order over point-major CUDA tensors:

    subsample(points, feat, targetnum, kp_idx)                     core/tf_utils.py:86-96
    feature_conv1d_1(feat, dim, name, c_last=True, ac_func=BNReLU)  core/tf_utils.py:99-109
    flexconv_withBatchnorm(feats, points, nn, dout, name)           core/tf_utils.py:48-64
    convolution_pointset_withBatchnorm(points, nn, dout, name)      core/tf_utils.py:67-83

The reference creates variables under the current TF variable scope; here the layer objects live in a
``layers.VariableStore`` under the same scope names ('<name>/tfconv0', '<name>', '<name>_bn').
"""
from . import ops
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID
from .backbones import FeatureConv1d, subsample  # noqa: F401  (subsample: same signature as the reference)
from .layers import BatchNorm, ConvolutionPointset, FlexConvolution, default_store

BNReLU = "BNReLU"   # the reference passes tensorpack's BNReLU / tf.nn.relu / tf.nn.sigmoid as ``ac_func``
_AC = {BNReLU: (True, ACT_RELU), "relu": (False, ACT_RELU), "sigmoid": (False, ACT_SIGMOID), None: (False, ACT_NONE)}


def feature_conv1d_1(feat, dim, name, c_last=True, ac_func=BNReLU, store=None):
    """feat [B,N,Cin] (c_last) or [B,Cin,N] -> [B,N,dim] / [B,dim,N]: tensorpack Conv2D(kernel 1) + ac_func."""
    bn, act = _AC[ac_func]
    st = store if store is not None else default_store
    x = feat if c_last else ops.transpose_cm_to_pm(feat.contiguous())
    layer = st.layer(name, lambda: FeatureConv1d(x.shape[-1], dim, bn=bn, act=act).to(feat.device))
    y = layer(x.contiguous())
    return y if c_last else ops.transpose_pm_to_cm(y)


def flexconv_withBatchnorm(feats, points, nn, dout, name, ac_func="relu", store=None):
    """feats [B,Din,N], points [B,3,N], nn [B,K,N] -> relu(BN(flexconv + feature_bias)) [B,dout,N]; the
    BatchNorm and the activation run in the FlexConv kernel's epilogue."""
    st = store if store is not None else default_store
    conv = st.layer(name, lambda: FlexConvolution(feats.shape[1], dout).to(feats.device))
    bn = st.layer(name + "_bn", lambda: BatchNorm(dout).to(feats.device))
    y = conv.forward_pm(ops.transpose_cm_to_pm(feats.contiguous()), ops.transpose_cm_to_pm(points.contiguous()),
                        ops.transpose_cm_to_pm(nn.contiguous()), bn=bn, act=_AC[ac_func][1])
    return ops.transpose_pm_to_cm(y)


def convolution_pointset_withBatchnorm(points, nn, dout, name, ac_func="relu", store=None):
    st = store if store is not None else default_store
    conv = st.layer(name, lambda: ConvolutionPointset(points.shape[1], dout).to(points.device))
    bn = st.layer(name + "_bn", lambda: BatchNorm(dout).to(points.device))
    y = conv.forward_pm(ops.transpose_cm_to_pm(points.contiguous()), ops.transpose_cm_to_pm(nn.contiguous()),
                        bn=bn, act=_AC[ac_func][1])
    return ops.transpose_pm_to_cm(y)
