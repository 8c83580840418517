"""ctypes binding of libdh3d_b200.so (the C ABI in include/dh3d_b200.h).

There is no CPU or PyTorch fallback anywhere in this package: if the shared library is missing or
a tensor is not a contiguous CUDA tensor of the right dtype the call raises.
"""
import ctypes
import os

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libdh3d_b200.so")

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2

_c_int, _c_float, _c_size_t, _p = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/dh3d_b200.h one to one
_SIGNATURES = {
    "dh3d_version": (_c_int, []),
    "dh3d_error_string": (ctypes.c_char_p, [_c_int]),
    "dh3d_knn_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "dh3d_knn_bruteforce": (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _c_size_t, _p]),
    "dh3d_knn_bruteforce_pm": (_c_int, [_p, _c_int, _c_int, _c_int, _p, _p, _p, _c_size_t, _p]),
    "dh3d_flex_conv_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "dh3d_flex_conv": (_c_int, [_p] * 6 + [_c_int] * 5 + [_p, _c_size_t, _p]),
    "dh3d_flex_conv_pm_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "dh3d_flex_conv_pm": (_c_int, [_p] * 6 + [_c_int] * 5 + [_p, _p, _p, _c_int, _p, _c_size_t, _p]),
    "dh3d_flex_conv_prepack_bytes": (_c_size_t, [_c_int] * 2),
    "dh3d_flex_conv_prepack": (_c_int, [_p] * 5 + [_c_int] * 2 + [_p, _p]),
    "dh3d_flex_conv_pm_packed_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "dh3d_flex_conv_pm_packed": (_c_int, [_p] * 5 + [_c_int] * 5 + [_p, _c_int, _p, _c_size_t, _p]),
    "dh3d_flex_pool": (_c_int, [_p] * 4 + [_c_int] * 4 + [_p]),
    "dh3d_flex_pool_pm": (_c_int, [_p] * 4 + [_c_int] * 4 + [_p]),
    "dh3d_conv_pointset": (_c_int, [_p] * 5 + [_c_int] * 5 + [_p]),
    "dh3d_conv_pointset_pm": (_c_int, [_p] * 5 + [_c_int] * 5 + [_p, _p, _c_int, _p]),
    "dh3d_farthest_point_sample": (_c_int, [_c_int, _c_int, _c_int, _p, _p, _p]),
    "dh3d_farthest_point_sample_presorted": (_c_int, [_c_int, _c_int, _c_int, _p, _p, _p]),
    "dh3d_knn_sort_pm": (_c_int, [_p, _c_int, _c_int, _p, _c_size_t, _p]),
    "dh3d_knn_query_sorted": (_c_int, [_p, _c_int, _c_int, _c_int, _p, _p, _p]),
    "dh3d_gather_point": (_c_int, [_c_int, _c_int, _c_int, _p, _p, _p, _p]),
    "dh3d_group_point": (_c_int, [_c_int] * 5 + [_p, _p, _p, _p]),
    "dh3d_query_ball_point_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "dh3d_query_ball_point": (_c_int, [_c_int, _c_int, _c_int, _c_float, _c_int, _p, _p, _p, _p, _p,
                                       _c_size_t, _p]),
    "dh3d_three_nn": (_c_int, [_c_int, _c_int, _c_int, _p, _p, _p, _p, _p]),
    "dh3d_three_nn_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "dh3d_three_nn_ws": (_c_int, [_c_int, _c_int, _c_int, _p, _p, _p, _p, _p, _c_size_t, _p]),
    "dh3d_three_nn_ws_presorted": (_c_int, [_c_int, _c_int, _c_int, _p, _p, _p, _p, _p, _c_size_t, _p]),
    "dh3d_three_nn_presorted2": (_c_int, [_c_int, _c_int, _c_int, _p, _p, _p, _p, _p]),
    "dh3d_three_interpolate": (_c_int, [_c_int] * 4 + [_p] * 5),
    "dh3d_three_interpolate_from_dist": (_c_int, [_c_int] * 4 + [_p] * 5),
    "dh3d_linear": (_c_int, [_p, _c_int, _p, _p, _p, _c_int, _p, _c_int, _c_int, _c_int, _c_int, _p]),
    "dh3d_linear_prepack_bytes": (_c_size_t, [_c_int, _c_int]),
    "dh3d_linear_prepack": (_c_int, [_p, _c_int, _c_int, _p, _p]),
    "dh3d_linear_packed": (_c_int, [_p, _c_int, _p, _p, _p, _c_int, _p, _c_int, _c_int, _c_int, _c_int, _p]),
    "dh3d_linear_rowdot_packed": (_c_int, [_p, _c_int, _p, _p, _p, _c_int, _p, _c_float, _c_int, _p, _c_int,
                                            _c_int, _c_int, _p]),
    "dh3d_rowdot": (_c_int, [_p, _c_int, _p, _c_float, _c_int, _p, _c_int, _c_int, _p]),
    "dh3d_linear_chain_packed": (_c_int, [_p, _c_int, _p, _p, _p, _c_int, _p, _p, _p, _c_int, _p, _c_int, _c_int, _c_int,
                                          _c_int, _c_int, _p]),
    "dh3d_linear_join_packed": (_c_int, [_p, _c_int, _p, _p, _p, _c_int, _p, _c_int, _p, _p, _p, _c_int, _p, _c_int,
                                         _p, _c_int, _c_float, _c_int, _c_int, _c_int, _c_int, _p]),
    "dh3d_se_excite": (_c_int, [_p, _p, _p, _c_size_t, _p]),
    "dh3d_se_pool_excite": (_c_int, [_p] * 7 + [_c_int] * 5 + [_p]),
    "dh3d_l2_normalize_rows": (_c_int, [_p, _c_int, _p, _c_int, _c_int, _c_int, _c_float, _p]),
    "dh3d_add": (_c_int, [_p, _p, _p, _c_size_t, _p]),
    "dh3d_copy_cols": (_c_int, [_p, _c_int, _p, _c_int, _c_int, _c_int, _p]),
    "dh3d_transpose_cm_to_pm": (_c_int, [_p, _p, _c_int, _c_int, _c_int, _p]),
    "dh3d_transpose_pm_to_cm": (_c_int, [_p, _p, _c_int, _c_int, _c_int, _p]),
    "dh3d_topk_l2": (_c_int, [_p, _c_int, _p, _p, _c_int, _c_int, _c_int, _p, _p, _p]),
    "dh3d_topk_l2_exact": (_c_int, [_p, _c_int, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _p, _p]),
    "dh3d_netvlad_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "dh3d_netvlad": (_c_int, [_p, _p] + [_c_int] * 5 + [_p] * 10 + [_c_int, _p, _p, _c_size_t, _p]),
    "dh3d_flex_conv_grad_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "dh3d_flex_conv_grad": (_c_int, [_p] * 9 + [_c_int] * 5 + [_p, _c_size_t, _p]),
    "dh3d_flex_conv_grad_pm_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "dh3d_flex_conv_grad_pm": (_c_int, [_p] * 9 + [_c_int] * 5 + [_p, _c_size_t, _p]),
    "dh3d_flex_pool_grad": (_c_int, [_p] * 3 + [_c_int] * 3 + [_p]),
    "dh3d_conv_pointset_grad_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "dh3d_conv_pointset_grad": (_c_int, [_p] * 7 + [_c_int] * 5 + [_p, _c_size_t, _p]),
    "dh3d_flex_deconv_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "dh3d_flex_deconv": (_c_int, [_p] * 6 + [_c_int] * 5 + [_p, _c_size_t, _p]),
    "dh3d_group_point_grad": (_c_int, [_c_int] * 5 + [_p, _p, _p, _p]),
    "dh3d_gather_point_grad": (_c_int, [_c_int] * 3 + [_p, _p, _p, _p]),
    "dh3d_three_interpolate_grad": (_c_int, [_c_int] * 4 + [_p] * 5),
    "dh3d_group_point_ld": (_c_int, [_c_int] * 5 + [_p, _c_int, _p, _p, _p]),
    "dh3d_three_interpolate_ld": (_c_int, [_c_int] * 4 + [_p, _p, _p, _c_int, _p, _c_int, _p]),
    "dh3d_add_l2_normalize_rows": (_c_int, [_p, _p, _p, _p, _c_int, _c_int, _c_float, _p]),
    "dh3d_affine": (_c_int, [_p, _c_float, _c_float, _p, _c_size_t, _p]),
    "dh3d_gather_rows": (_c_int, [_c_int] * 4 + [_p, _p, _p, _c_int, _p]),
    "dh3d_keypoint_nms_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "dh3d_keypoint_nms": (_c_int, [_p, _p, _c_int, _c_int, _c_float, _c_float, _c_int, _c_int, _p, _p, _p,
                                   _c_size_t, _p]),
}

_lib = None

# kernels launched per C-ABI call (memcpy/memset nodes not counted); bench.py's gpu_launches claim
_KERNELS_PER_CALL = {
    "dh3d_knn_bruteforce": 2, "dh3d_knn_bruteforce_pm": 2,          # pack + scan
    "dh3d_flex_conv": 8,                                             # 4 transposes + 2 theta_ext + fold_bias + fused kernel
    "dh3d_flex_conv_pm": 4,                                          # 2 theta_ext forms + fold_bias + the fused kernel
    "dh3d_flex_conv_prepack": 3, "dh3d_flex_conv_pm_packed": 1,      # weights once; then the fused kernel only
    "dh3d_query_ball_point": 2, "dh3d_netvlad": 3, "dh3d_three_nn_ws": 3, "dh3d_three_nn_ws_presorted": 2,   # (knn_sort_pm / knn_query_sorted / fps_presorted: 1 each) netvlad: cluster-weight prepack + aggregate + tail
    "dh3d_flex_conv_grad_pm": 6, "dh3d_flex_conv_grad": 11, "dh3d_conv_pointset_grad": 5, "dh3d_flex_deconv": 7,
    "dh3d_keypoint_nms": 5,
}


class Stats(object):
    """Launch accounting and optional per-op device timing (CUDA events on the launching stream)."""

    def __init__(self):
        self.calls = 0
        self.kernels = 0
        self.timing_filter = None   # None = off; set() of entry-point names, or "all"
        self.events = []            # (name, start_event, end_event, tag)
        self.tag = None

    def reset(self):
        self.calls = self.kernels = 0
        self.events = []

    def op_times_ms(self):
        """Aggregate recorded events -> {name or (name, tag): [total_ms, calls]} (call after sync)."""
        agg = {}
        for name, a, b, tag in self.events:
            key = name if tag is None else "%s[%s]" % (name, tag)
            t = agg.setdefault(key, [0.0, 0])
            t[0] += a.elapsed_time(b)
            t[1] += 1
        return agg


stats = Stats()


class Dh3dError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Dh3dError(
                "%s is missing -- build it with `python -m dh3d_b200.build` (or "
                "__graft_entry__.build()).  There is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def error_string(code):
    return lib().dh3d_error_string(int(code)).decode()


def call(name, *args):
    fn = getattr(lib(), name)
    tf = stats.timing_filter
    if tf is not None and (tf == "all" or name in tf):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        stats.events.append((name, a, b, stats.tag))
    else:
        rc = fn(*args)
    stats.calls += 1
    stats.kernels += _KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise Dh3dError("%s failed with code %d: %s" % (name, rc, error_string(rc)))


def query(name, *args):
    return int(getattr(lib(), name)(*args))


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def check(t, dtype, name, ndim=None):
    """Validate a tensor for the C ABI: CUDA, contiguous, exact dtype.  Returns its data pointer."""
    if not isinstance(t, torch.Tensor):
        raise Dh3dError("%s: expected a torch.Tensor, got %r" % (name, type(t)))
    if not t.is_cuda:
        raise Dh3dError("%s: expected a CUDA tensor (no CPU path exists), got device %s" % (name, t.device))
    if t.dtype != dtype:
        raise Dh3dError("%s: expected dtype %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise Dh3dError("%s: expected a contiguous tensor" % name)
    if ndim is not None and t.dim() != ndim:
        raise Dh3dError("%s: expected rank %d, got shape %s" % (name, ndim, tuple(t.shape)))
    return ctypes.c_void_p(t.data_ptr())


def opt(t, dtype, name):
    return ctypes.c_void_p(0) if t is None else check(t, dtype, name)


def workspace(nbytes, device):
    """Caller-owned scratch for one call (torch's caching allocator makes this cheap and
    stream-safe; 512-byte aligned)."""
    t = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
    return t, ctypes.c_void_p(t.data_ptr()), ctypes.c_size_t(t.numel())
