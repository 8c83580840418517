"""Read-only loader for the reference's TF1 checkpoints -- no TensorFlow.

A TensorBundle is `<prefix>.index` (a LevelDB-format SSTable: key = variable name, value = a
BundleEntryProto with dtype / shape / shard / offset / size) plus `<prefix>.data-00000-of-00001`
(raw little-endian tensor bytes).  Both formats are restated from their public specifications
(leveldb `table_format.md`; tensorflow/core/protobuf/tensor_bundle.proto field numbers).

    tensors = read_tensor_bundle("/root/reference/models/global/globalmodel")   # {tf_name: ndarray}
    load_reference_checkpoint(model, local_prefix, global_prefix)               # fills DH3D modules

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
                 "local_stage1_shortcut")
_NETVLAD_ROOT = ("cluster_weights", "cluster_weights2", "cluster_bn", "hidden1_weights", "bn",
                 "gating_weights", "gating_bn")


def tf_name_to_param(name):
    """'stage1/flexconv_0_bn/mean/EMA' -> 'local.stage1.flexconv_0_bn.mean_ema' (None = not a model
    variable: optimizer slots, global_step, ...)."""
    if "Adam" in name or name in ("global_step", "learning_rate") or name.startswith("beta"):
        return None
    name = name.replace("mean/EMA", "mean_ema").replace("variance/EMA", "variance_ema")
    name = name.replace("moving_mean", "mean_ema").replace("moving_variance", "variance_ema")
    parts = name.split("/")
    if parts[0] in _LOCAL_SCOPES:
        parts = ["local"] + parts
    elif parts[0] in _NETVLAD_ROOT:
        parts = ["netvlad"] + parts
    return ".".join(parts)


def load_tensors_into(model, tensors, strict_shapes=True):
    """Copy {tf_name: ndarray} into the model's same-named parameters; returns (loaded, missing)."""
    import torch
    from .layers import invalidate_folded
    params = dict(model.named_parameters())
    loaded = set()
    with torch.no_grad():
        for tf_name, arr in tensors.items():
            pname = tf_name_to_param(tf_name)
            if pname is None or pname not in params:
                continue
            p = params[pname]
            if tuple(p.shape) != tuple(arr.shape):
                if strict_shapes:
                    raise ValueError("%s: checkpoint shape %s != parameter shape %s" %
                                     (tf_name, arr.shape, tuple(p.shape)))
                continue
            p.copy_(torch.from_numpy(np.array(arr)).to(p.dtype))
            loaded.add(pname)
    invalidate_folded(model)
    return sorted(loaded), sorted(set(params) - loaded)


def load_reference_checkpoint(model, local_prefix=None, global_prefix=None):
    """Fill a DH3D model from the shipped checkpoints: the local checkpoint provides the backbone and
    the detector, the global one the global branch (and its own copy of the backbone, which wins for
    the shared variables when both are given -- the two are NOT identical)."""
    loaded = []
    if local_prefix:
        loaded += load_tensors_into(model, read_tensor_bundle(local_prefix))[0]
    if global_prefix:
        loaded += load_tensors_into(model, read_tensor_bundle(global_prefix))[0]
    missing = sorted(set(dict(model.named_parameters())) - set(loaded))
    return sorted(set(loaded)), missing
