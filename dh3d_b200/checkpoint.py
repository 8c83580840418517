"""Read-only loader for the reference's TF1 checkpoints -- no TensorFlow.

A TensorBundle is `<prefix>.index` (a LevelDB-format SSTable: key = variable name, value = a
BundleEntryProto with dtype / shape / shard / offset / size) plus `<prefix>.data-00000-of-00001`
(raw little-endian tensor bytes).  Both formats are restated from their public specifications
(leveldb `table_format.md`; tensorflow/core/protobuf/tensor_bundle.proto field numbers).

    tensors = read_tensor_bundle("/root/reference/models/global/globalmodel")   # {tf_name: ndarray}
    load_reference_checkpoint(model, local_prefix, global_prefix)               # fills DH3D modules
    checkpoint_model(local_prefix, global_prefix)                               # builds + fills the right DH3D

Name mapping tf -> dh3d_b200.model.DH3D: '/' -> '.', tensorpack 'mean/EMA','variance/EMA' and slim
'moving_mean','moving_variance' -> 'mean_ema','variance_ema'; backbone scopes get the 'local.'
prefix and the root-scope NetVLAD variables the 'netvlad.' prefix.
"""
import struct

import numpy as np

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}
_MAGIC = 0xDB4775248B80FB57


def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(buf, offset, size):
    """Yield (key, value) of one SSTable block (prefix-compressed keys + restart array)."""
    block = buf[offset:offset + size]
    if buf[offset + size] != 0:
        raise ValueError("compressed SSTable blocks are not supported")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _parse_proto(buf):
    """Minimal protobuf wire parser -> {field: [values]} (varint, 64-bit, length-delimited, 32-bit)."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _varint(buf, pos)
        elif wire == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wire == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.setdefault(field, []).append(val)
    return out


def read_tensor_bundle(prefix):
    index = open(prefix + ".index", "rb").read()
    if struct.unpack_from("<Q", index, len(index) - 8)[0] != _MAGIC:
        raise ValueError("%s.index is not an SSTable" % prefix)
    footer = index[len(index) - 48:]
    _, p = _varint(footer, 0)          # metaindex handle (offset, size) -- unused
    _, p = _varint(footer, p)
    idx_off, p = _varint(footer, p)
    idx_size, p = _varint(footer, p)
    data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
    tensors = {}
    for _, handle in _block_entries(index, idx_off, idx_size):
        boff, q = _varint(handle, 0)
        bsize, q = _varint(handle, q)
        for key, value in _block_entries(index, boff, bsize):
            if not key:
                continue  # BundleHeaderProto
            e = _parse_proto(value)
            dtype = _DTYPES.get(e.get(1, [0])[0])
            if dtype is None:
                continue
            shape = []
            if 2 in e:
                for dim in _parse_proto(e[2][0]).get(2, []):
                    shape.append(_parse_proto(dim).get(1, [0])[0])
            off, size = e.get(4, [0])[0], e.get(5, [0])[0]
            arr = np.frombuffer(bytes(data[off:off + size]), dtype=dtype).reshape(shape)
            tensors[key.decode()] = arr
    return tensors


_LOCAL_SCOPES = ("initconv", "initconv_bn", "stage1", "before_stage2_conv1d", "stage2",
                 "local_stage1_shortcut", "final_fc")
_NETVLAD_ROOT = ("cluster_weights", "cluster_weights2", "cluster_bn", "hidden1_weights", "bn",
                 "gating_weights", "gating_bn")


class CheckpointError(ValueError):
    pass


def tf_name_to_param(name, backbone="local"):
    """'stage1/flexconv_0_bn/mean/EMA' -> 'local.stage1.flexconv_0_bn.mean_ema'.  ``backbone`` is the module
    the backbone scopes map to ('local', or 'global_local' for the global checkpoint's own copy).  Purely
    syntactic: whether the result names a model variable is decided by the caller against the model itself
    (optimizer slots such as '.../W/Adam_1', 'global_step', 'beta1_power' simply map to names no model has)."""
    name = name.replace("mean/EMA", "mean_ema").replace("variance/EMA", "variance_ema")
    name = name.replace("moving_mean", "mean_ema").replace("moving_variance", "variance_ema")
    parts = name.split("/")
    if parts[0] in _LOCAL_SCOPES:
        parts = [backbone] + parts
    elif parts[0] in _NETVLAD_ROOT:
        parts = ["netvlad"] + parts
    return ".".join(parts)


def load_tensors_into(model, tensors, strict_shapes=True, backbone="local", only=None):
    """Copy {tf_name: ndarray} into the model's same-named parameters; returns (loaded, missing).
    ``only``: restrict to parameters whose dotted name starts with one of these prefixes."""
    import torch
    from .layers import invalidate_folded
    params = dict(model.named_parameters())
    loaded = set()
    with torch.no_grad():
        for tf_name, arr in tensors.items():
            pname = tf_name_to_param(tf_name, backbone)
            if pname not in params:          # optimizer slots, summaries, branches this model does not have
                continue
            if only is not None and not pname.startswith(tuple(only)):
                continue
            p = params[pname]
            if tuple(p.shape) != tuple(arr.shape):
                if strict_shapes:
                    raise CheckpointError("%s: checkpoint shape %s != parameter shape %s" %
                                          (tf_name, arr.shape, tuple(p.shape)))
                continue
            p.copy_(torch.from_numpy(np.array(arr)).to(p.dtype))
            loaded.add(pname)
    invalidate_folded(model)
    return sorted(loaded), sorted(set(params) - loaded)


def _shared_backbone_differs(a, b, tol=0.0):
    """Names of backbone variables present in both bundles whose values differ."""
    out = []
    for k, v in a.items():
        if k.split("/")[0] in _LOCAL_SCOPES and k in b and "Adam" not in k:
            if v.shape != b[k].shape or float(np.abs(v - b[k]).max()) > tol:
                out.append(k)
    return out


def load_reference_checkpoint(model, local_prefix=None, global_prefix=None, strict=True, shared_backbone=None):
    """Fill a DH3D model from the reference's checkpoints (either may also be a {tf_name: ndarray} dict).

    The local checkpoint provides the backbone + detector, the global one the global branch AND its own,
    differently trained copy of the backbone (the shipped pair differs by up to 0.12 in the stage-1
    FlexConv weights).  When both are given:
      * a model built with ``separate_global_backbone=True`` gets each backbone from its own checkpoint
        (``local.*`` <- local, ``global_local.*`` <- global): both reference networks reproduced in one pass;
      * a model with ONE backbone must be told which copy it gets, ``shared_backbone='local'|'global'``, if the
        two differ -- otherwise this raises instead of silently evaluating one head on the other's backbone.
    ``strict`` (default): raise CheckpointError if any parameter of the model was left unfilled."""
    lt = local_prefix if isinstance(local_prefix, dict) or local_prefix is None else read_tensor_bundle(local_prefix)
    gt = global_prefix if isinstance(global_prefix, dict) or global_prefix is None else read_tensor_bundle(global_prefix)
    separate = getattr(model, "global_local", None) is not None
    loaded = set()
    if lt is not None and gt is not None and not separate:
        differ = _shared_backbone_differs(lt, gt)
        if differ and shared_backbone not in ("local", "global"):
            raise CheckpointError(
                "the two checkpoints carry different backbones (%d variables differ, e.g. %s): build the model with "
                "separate_global_backbone=True, or pass shared_backbone='local'|'global' to choose one"
                % (len(differ), differ[0]))
    if lt is not None:
        skip_backbone = (gt is not None and not separate and shared_backbone == "global")
        only = None if not skip_backbone else ("detection_block_reliable",)
        loaded |= set(load_tensors_into(model, lt, only=only)[0])
    if gt is not None:
        if separate:
            loaded |= set(load_tensors_into(model, gt, backbone="global_local")[0])
        else:
            only = None
            if lt is not None and shared_backbone == "local":
                only = ("global_before_assemble", "globalatt", "netvlad")
            loaded |= set(load_tensors_into(model, gt, only=only)[0])
    missing = sorted(set(dict(model.named_parameters())) - loaded)
    if strict and missing:
        raise CheckpointError("%d model parameters are in neither checkpoint (e.g. %s): the model enables a branch "
                              "the checkpoint(s) do not carry" % (len(missing), ", ".join(missing[:4])))
    return sorted(loaded), missing


def checkpoint_model(local_prefix=None, global_prefix=None, **config_overrides):
    """Build the DH3D model the given checkpoint(s) describe and fill it: local only -> detection_config,
    global only -> global_config, both -> full_config with separate backbones."""
    from .configs import DH3DConfig
    from .model import DH3D
    cfg = DH3DConfig(detection=local_prefix is not None, extract_global=global_prefix is not None, **config_overrides)
    model = DH3D(cfg, separate_global_backbone=local_prefix is not None and global_prefix is not None)
    load_reference_checkpoint(model, local_prefix, global_prefix)
    return model
