"""CPU oracle for the DH3D hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the CPU
comparator.  The product (``dh3d_b200``) never imports it.

* ``dh3d_oracle.c``  -- plain-C restatement of every custom op (each function cites the
  reference file:line it follows); wrapped here with ctypes over numpy arrays.
* ``net.py``         -- numpy fp64 restatement of the TF-library part of the forward
  (1x1 convs, BN, SE, detector, attention, NetVLAD; reference ``core/backbones.py``).
* ``fixtures.py``    -- the reference's own test fixtures, regenerated (seed-42
  ``FakePointCloud``, ``python_bruteforce``, the 4-point FlexPool case).
* ``_ref/``          -- (git-ignored) the reference's own CUDA / C++ sources compiled
  unmodified from /root/reference by ``build_ref.py``; used to pin this oracle.

Parity pinning status: see ``oracle/README.md``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdh3d_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile dh3d_oracle.c -> libdh3d_oracle.so (gcc, seconds)."""
    src = os.path.join(_HERE, "dh3d_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", _HERE, "-B", "libdh3d_oracle.so"], check=True, env=env,
                   stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    if a.dtype == np.float32:
        return a.ctypes.data_as(_f32p)
    if a.dtype == np.float64:
        return a.ctypes.data_as(_f64p)
    if a.dtype == np.int32:
        return a.ctypes.data_as(_i32p)
    raise TypeError(a.dtype)


def knn_tile(N):
    T, V = ctypes.c_int(), ctypes.c_int()
    lib().orc_knn_tile(int(N), ctypes.byref(T), ctypes.byref(V))
    return T.value, V.value


def knn_bruteforce(positions, k, literal=False):
    """positions [B,Dp,N] -> (ids [B,N,K] i32, dists [B,N,K] f32); reference CUDA tie order."""
    pos = _f32(positions)
    B, Dp, N = pos.shape
    ids = np.empty((B, N, k), np.int32)
    dist = np.empty((B, N, k), np.float32)
    fn = lib().orc_knn_literal if literal else lib().orc_knn
    fn(B, Dp, N, int(k), _p(pos), _p(ids), _p(dist))
    return ids, dist


def flex_convolution(features, position, neighborhood, theta, bias, centre_is_self=True,
                     f64=False):
    """Argument order of user_ops.flex_convolution (user_ops/__init__.py:63-89)."""
    f, p, nb = _f32(features), _f32(position), _i32(neighborhood)
    th, bi = _f32(theta), _f32(bias)
    B, Din, N = f.shape
    K = nb.shape[1]
    Dout = th.shape[2]
    assert th.shape == (3, Din, Dout) and bi.shape == (Din, Dout) and p.shape == (B, 3, N)
    if f64:
        out = np.empty((B, Dout, N), np.float64)
        lib().orc_flex_conv_f64(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(p),
                                _p(out), int(centre_is_self))
    else:
        out = np.empty((B, Dout, N), np.float32)
        lib().orc_flex_conv(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(p), _p(out),
                            int(centre_is_self))
    return out


def flex_pooling(features, neighborhood):
    f, nb = _f32(features), _i32(neighborhood)
    B, D, N = f.shape
    K = nb.shape[1]
    out = np.empty((B, D, N), np.float32)
    arg = np.empty((B, D, N), np.int32)
    lib().orc_flex_pool(B, N, K, D, _p(f), _p(nb), _p(out), _p(arg))
    return out, arg


def convolution_pointset(features, neighborhood, theta, bias):
    f, nb, th, bi = _f32(features), _i32(neighborhood), _f32(theta), _f32(bias)
    B, Din, N = f.shape
    K = nb.shape[1]
    Dout = th.shape[1]
    assert th.shape == (Din, Dout) and bi.shape == (Dout,)
    out = np.empty((B, Dout, N), np.float32)
    lib().orc_conv_pointset(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(out))
    return out


def farthest_point_sample(npoint, inp):
    x = _f32(inp)
    B, N, _ = x.shape
    idx = np.empty((B, npoint), np.int32)
    lib().orc_fps(B, N, int(npoint), _p(x), _p(idx))
    return idx


def gather_point(inp, idx):
    x, i = _f32(inp), _i32(idx)
    B, N, _ = x.shape
    M = i.shape[1]
    out = np.empty((B, M, 3), np.float32)
    lib().orc_gather_point(B, N, M, _p(x), _p(i), _p(out))
    return out


def group_point(points, idx):
    x, i = _f32(points), _i32(idx)
    B, N, C = x.shape
    _, M, S = i.shape
    out = np.empty((B, M, S, C), np.float32)
    lib().orc_group_point(B, N, C, M, S, _p(x), _p(i), _p(out))
    return out


def query_ball_point(radius, nsample, xyz1, xyz2):
    a, b = _f32(xyz1), _f32(xyz2)
    B, n, _ = a.shape
    m = b.shape[1]
    idx = np.zeros((B, m, nsample), np.int32)
    cnt = np.zeros((B, m), np.int32)
    lib().orc_query_ball_point(B, n, m, ctypes.c_float(radius), int(nsample), _p(a), _p(b),
                               _p(idx), _p(cnt))
    return idx, cnt


def three_nn(xyz1, xyz2):
    a, b = _f32(xyz1), _f32(xyz2)
    B, n, _ = a.shape
    m = b.shape[1]
    dist = np.empty((B, n, 3), np.float32)
    idx = np.empty((B, n, 3), np.int32)
    lib().orc_three_nn(B, n, m, _p(a), _p(b), _p(dist), _p(idx))
    return dist, idx


def three_interpolate(points, idx, weight):
    p, i, w = _f32(points), _i32(idx), _f32(weight)
    B, m, c = p.shape
    n = i.shape[1]
    out = np.empty((B, n, c), np.float32)
    lib().orc_three_interpolate(B, m, c, n, _p(p), _p(i), _p(w), _p(out))
    return out


def three_nn_weights(dist):
    """backbones.py:92-95: d=max(d,1e-10); w=(1/d)/sum(1/d) in fp32 (TF elementwise ops)."""
    d = np.maximum(np.asarray(dist, np.float32), np.float32(1e-10))
    inv = (np.float32(1.0) / d).astype(np.float32)
    norm = inv[..., 0:1] + inv[..., 1:2]
    norm = (norm + inv[..., 2:3]).astype(np.float32)
    return (inv / norm).astype(np.float32)


# ---- backward passes / FlexDeconv: fp64 truth of the reference's CPU loops ---------------------
def flex_convolution_grad(features, theta, bias, neighborhood, position, topdiff):
    """Op-input order of FlexConvGrad (user_ops/__init__.py:95-111) -> (df, dtheta, dbias) fp64."""
    f, th, bi = _f32(features), _f32(theta), _f32(bias)
    nb, p, t = _i32(neighborhood), _f32(position), _f32(topdiff)
    B, Din, N = f.shape
    K, Dout = nb.shape[1], th.shape[2]
    gf = np.empty((B, Din, N), np.float64)
    gt = np.empty((3, Din, Dout), np.float64)
    gb = np.empty((Din, Dout), np.float64)
    lib().orc_flex_conv_grad(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(p), _p(t),
                             _p(gf), _p(gt), _p(gb))
    return gf, gt, gb


def flex_pooling_grad(topdiff, argmax):
    t, a = _f32(topdiff), _i32(argmax)
    B, D, N = t.shape
    gf = np.empty((B, D, N), np.float64)
    lib().orc_flex_pool_grad(B, N, D, _p(t), _p(a), _p(gf))
    return gf


def convolution_pointset_grad(features, theta, neighborhood, topdiff):
    f, th, nb, t = _f32(features), _f32(theta), _i32(neighborhood), _f32(topdiff)
    B, Din, N = f.shape
    K, Dout = nb.shape[1], th.shape[1]
    gf = np.empty((B, Din, N), np.float64)
    gt = np.empty((Din, Dout), np.float64)
    gb = np.empty((Dout,), np.float64)
    lib().orc_conv_pointset_grad(B, N, K, Din, Dout, _p(f), _p(th), _p(nb), _p(t), _p(gf), _p(gt), _p(gb))
    return gf, gt, gb


def flex_convolution_transpose(features, position, neighborhood, theta, bias):
    f, p, nb, th, bi = _f32(features), _f32(position), _i32(neighborhood), _f32(theta), _f32(bias)
    B, Din, N = f.shape
    K, Dout = nb.shape[1], th.shape[2]
    out = np.empty((B, Dout, N), np.float64)
    lib().orc_flex_deconv(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(p), _p(out))
    return out


def group_point_grad(n, grad_out, idx):
    g, i = _f32(grad_out), _i32(idx)
    B, M, S, C = g.shape
    gp = np.empty((B, n, C), np.float64)
    lib().orc_group_point_grad(B, int(n), C, M, S, _p(g), _p(i), _p(gp))
    return gp


def three_interpolate_grad(m, grad_out, idx, weight):
    g, i, w = _f32(grad_out), _i32(idx), _f32(weight)
    B, N, C = g.shape
    gp = np.empty((B, int(m), C), np.float64)
    lib().orc_three_interpolate_grad(B, N, C, int(m), _p(g), _p(i), _p(w), _p(gp))
    return gp
