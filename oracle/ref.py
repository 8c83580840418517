"""ctypes access to oracle/_ref/*.so -- the reference's OWN code compiled unmodified by
oracle/build_ref.py (test / bench infrastructure only).  CPU entry points take numpy arrays; CUDA
entry points take CUDA torch tensors and run the reference kernels on the default stream."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_SO = os.path.join(_HERE, "_ref", "libdh3d_ref_cuda.so")
CPU_SO = os.path.join(_HERE, "_ref", "libdh3d_ref_cpu.so")
_cuda = _cpu = None


def have_cpu():
    return os.path.exists(CPU_SO)


def have_cuda():
    return os.path.exists(CUDA_SO)


def cpu():
    global _cpu
    if _cpu is None:
        _cpu = ctypes.CDLL(CPU_SO)
    return _cpu


def cuda():
    global _cuda
    if _cuda is None:
        _cuda = ctypes.CDLL(CUDA_SO)
    return _cuda


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


# ---- reference CPU code ----------------------------------------------------------------------
def cpu_three_nn(xyz1, xyz2):
    a, b = _np(xyz1, np.float32), _np(xyz2, np.float32)
    B, n, _ = a.shape
    m = b.shape[1]
    dist, idx = np.empty((B, n, 3), np.float32), np.empty((B, n, 3), np.int32)
    cpu().ref_cpu_three_nn(B, n, m, _p(a), _p(b), _p(dist), _p(idx))
    return dist, idx


def cpu_three_interpolate(points, idx, weight):
    p, i, w = _np(points, np.float32), _np(idx, np.int32), _np(weight, np.float32)
    B, m, c = p.shape
    n = i.shape[1]
    out = np.empty((B, n, c), np.float32)
    cpu().ref_cpu_three_interpolate(B, m, c, n, _p(p), _p(i), _p(w), _p(out))
    return out


def cpu_flex_conv(features, position, neighborhood, theta, bias):
    f, p, nb = _np(features, np.float32), _np(position, np.float32), _np(neighborhood, np.int32)
    th, bi = _np(theta, np.float32), _np(bias, np.float32)
    B, Din, N = f.shape
    K, Dout = nb.shape[1], th.shape[2]
    out = np.empty((B, Dout, N), np.float32)
    cpu().ref_cpu_flex_conv(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(p), _p(out))
    return out


def cpu_flex_pool(features, neighborhood):
    f, nb = _np(features, np.float32), _np(neighborhood, np.int32)
    B, D, N = f.shape
    K = nb.shape[1]
    out, arg = np.empty((B, D, N), np.float32), np.empty((B, D, N), np.int32)
    cpu().ref_cpu_flex_pool(B, N, K, D, _p(f), _p(nb), _p(out), _p(arg))
    return out, arg


def cpu_conv_pointset(features, neighborhood, theta, bias):
    f, nb = _np(features, np.float32), _np(neighborhood, np.int32)
    th, bi = _np(theta, np.float32), _np(bias, np.float32)
    B, Din, N = f.shape
    K, Dout = nb.shape[1], th.shape[1]
    out = np.empty((B, Dout, N), np.float32)
    cpu().ref_cpu_conv_pointset(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(out))
    return out


# ---- reference CUDA kernels (torch CUDA tensors in, torch CUDA tensors out) -------------------
def _d(t):
    return ctypes.c_void_p(t.data_ptr())


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError("reference CUDA %s failed: %d" % (what, rc))


def cuda_knn(positions, k):
    import torch
    B, Dp, N = positions.shape
    ids = torch.empty((B, N, k), dtype=torch.int32, device=positions.device)
    dist = torch.empty((B, N, k), dtype=torch.float32, device=positions.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_knn(B, Dp, N, int(k), _d(positions.contiguous()), _d(ids), _d(dist)), "knn")
    return ids, dist


def cuda_fps(npoint, inp):
    import torch
    B, N, _ = inp.shape
    out = torch.empty((B, npoint), dtype=torch.int32, device=inp.device)
    temp = torch.empty((32, N), dtype=torch.float32, device=inp.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_fps(B, N, int(npoint), _d(inp.contiguous()), _d(temp), _d(out)), "fps")
    return out


def cuda_gather_point(inp, idx):
    import torch
    B, N, _ = inp.shape
    M = idx.shape[1]
    out = torch.empty((B, M, 3), dtype=torch.float32, device=inp.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_gather_point(B, N, M, _d(inp.contiguous()), _d(idx.contiguous()), _d(out)), "gather")
    return out


def cuda_group_point(points, idx):
    import torch
    B, N, C = points.shape
    _, M, S = idx.shape
    out = torch.empty((B, M, S, C), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_group_point(B, N, C, M, S, _d(points.contiguous()), _d(idx.contiguous()), _d(out)), "group")
    return out


def cuda_query_ball_point(radius, nsample, xyz1, xyz2):
    import torch
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = torch.zeros((B, m, nsample), dtype=torch.int32, device=xyz1.device)
    cnt = torch.zeros((B, m), dtype=torch.int32, device=xyz1.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_query_ball_point(B, n, m, ctypes.c_float(radius), int(nsample), _d(xyz1.contiguous()),
                                     _d(xyz2.contiguous()), _d(idx), _d(cnt)), "ball query")
    return idx, cnt


def cuda_flex_conv(features, position, neighborhood, theta, bias):
    import torch
    B, Din, N = features.shape
    K, Dout = neighborhood.shape[1], theta.shape[2]
    out = torch.empty((B, Dout, N), dtype=torch.float32, device=features.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_flex_conv(B, N, K, Din, Dout, _d(features.contiguous()), _d(theta.contiguous()),
                              _d(bias.contiguous()), _d(neighborhood.contiguous()),
                              _d(position.contiguous()), _d(out)), "flex_conv")
    return out


def cuda_flex_pool(features, neighborhood):
    import torch
    B, D, N = features.shape
    K = neighborhood.shape[1]
    out = torch.empty_like(features)
    arg = torch.empty((B, D, N), dtype=torch.int32, device=features.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_flex_pool(B, N, K, D, _d(features.contiguous()), _d(neighborhood.contiguous()), _d(out),
                              _d(arg)), "flex_pool")
    return out, arg


def cuda_conv_pointset(features, neighborhood, theta, bias):
    import torch
    B, Din, N = features.shape
    K, Dout = neighborhood.shape[1], theta.shape[1]
    out = torch.empty((B, Dout, N), dtype=torch.float32, device=features.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_conv_pointset(B, N, K, Din, Dout, _d(features.contiguous()), _d(theta.contiguous()),
                                  _d(bias.contiguous()), _d(neighborhood.contiguous()), _d(out)), "conv_pointset")
    return out


# ---- backward passes / FlexDeconv: reference CPU functors ----------------------------------------
def cpu_flex_conv_grad(features, theta, bias, neighborhood, position, topdiff):
    f, th, bi = _np(features, np.float32), _np(theta, np.float32), _np(bias, np.float32)
    nb, p, t = _np(neighborhood, np.int32), _np(position, np.float32), _np(topdiff, np.float32)
    B, Din, N = f.shape
    K, Dout = nb.shape[1], th.shape[2]
    gf, gt, gb = np.empty((B, Din, N), np.float32), np.empty((3, Din, Dout), np.float32), np.empty((Din, Dout), np.float32)
    cpu().ref_cpu_flex_conv_grad(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(p), _p(t), _p(gf), _p(gt), _p(gb))
    return gf, gt, gb


def cpu_flex_pool_grad(features, neighborhood, topdiff, argmax):
    f, nb = _np(features, np.float32), _np(neighborhood, np.int32)
    t, a = _np(topdiff, np.float32), _np(argmax, np.int32)
    B, D, N = f.shape
    gf = np.empty((B, D, N), np.float32)
    cpu().ref_cpu_flex_pool_grad(B, N, nb.shape[1], D, _p(f), _p(nb), _p(t), _p(a), _p(gf))
    return gf


def cpu_conv_pointset_grad(features, theta, bias, neighborhood, topdiff):
    f, th, bi = _np(features, np.float32), _np(theta, np.float32), _np(bias, np.float32)
    nb, t = _np(neighborhood, np.int32), _np(topdiff, np.float32)
    B, Din, N = f.shape
    K, Dout = nb.shape[1], th.shape[1]
    gf, gt, gb = np.empty((B, Din, N), np.float32), np.empty((Din, Dout), np.float32), np.empty((Dout,), np.float32)
    cpu().ref_cpu_conv_pointset_grad(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(t), _p(gf), _p(gt), _p(gb))
    return gf, gt, gb


def cpu_flex_deconv(features, position, neighborhood, theta, bias):
    f, p, nb = _np(features, np.float32), _np(position, np.float32), _np(neighborhood, np.int32)
    th, bi = _np(theta, np.float32), _np(bias, np.float32)
    B, Din, N = f.shape
    K, Dout = nb.shape[1], th.shape[2]
    out = np.empty((B, Dout, N), np.float32)
    cpu().ref_cpu_flex_deconv(B, N, K, Din, Dout, _p(f), _p(th), _p(bi), _p(nb), _p(p), _p(out))
    return out


def cpu_three_interpolate_grad(m, grad_out, idx, weight):
    g, i, w = _np(grad_out, np.float32), _np(idx, np.int32), _np(weight, np.float32)
    B, n, c = g.shape
    gp = np.zeros((B, int(m), c), np.float32)
    cpu().ref_cpu_three_interpolate_grad(B, n, c, int(m), _p(g), _p(i), _p(w), _p(gp))
    return gp


# ---- backward passes / FlexDeconv: reference CUDA kernels ----------------------------------------
def cuda_flex_conv_grad(features, theta, bias, neighborhood, position, topdiff):
    import torch
    B, Din, N = features.shape
    K, Dout = neighborhood.shape[1], theta.shape[2]
    dev = features.device
    gf = torch.empty((B, Din, N), dtype=torch.float32, device=dev)
    gt = torch.empty((3, Din, Dout), dtype=torch.float32, device=dev)
    gb = torch.empty((Din, Dout), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    _chk(cuda().ref_flex_conv_grad(B, N, K, Din, Dout, _d(features.contiguous()), _d(theta.contiguous()),
                                   _d(bias.contiguous()), _d(neighborhood.contiguous()), _d(position.contiguous()),
                                   _d(topdiff.contiguous()), _d(gf), _d(gt), _d(gb)), "flex_conv_grad")
    return gf, gt, gb


def cuda_flex_pool_grad(features, neighborhood, topdiff, argmax):
    import torch
    B, D, N = features.shape
    gf = torch.empty((B, D, N), dtype=torch.float32, device=features.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_flex_pool_grad(B, N, neighborhood.shape[1], D, _d(features.contiguous()),
                                   _d(neighborhood.contiguous()), _d(topdiff.contiguous()), _d(argmax.contiguous()),
                                   _d(gf)), "flex_pool_grad")
    return gf


def cuda_conv_pointset_grad(features, theta, bias, neighborhood, topdiff):
    import torch
    B, Din, N = features.shape
    K, Dout = neighborhood.shape[1], theta.shape[1]
    dev = features.device
    gf = torch.empty((B, Din, N), dtype=torch.float32, device=dev)
    gt = torch.empty((Din, Dout), dtype=torch.float32, device=dev)
    gb = torch.empty((Dout,), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    _chk(cuda().ref_conv_pointset_grad(B, N, K, Din, Dout, _d(features.contiguous()), _d(theta.contiguous()),
                                       _d(bias.contiguous()), _d(neighborhood.contiguous()), _d(topdiff.contiguous()),
                                       _d(gf), _d(gt), _d(gb)), "conv_pointset_grad")
    return gf, gt, gb


def cuda_flex_deconv(features, position, neighborhood, theta, bias):
    import torch
    B, Din, N = features.shape
    K, Dout = neighborhood.shape[1], theta.shape[2]
    out = torch.empty((B, Dout, N), dtype=torch.float32, device=features.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_flex_deconv(B, N, K, Din, Dout, _d(features.contiguous()), _d(theta.contiguous()),
                                _d(bias.contiguous()), _d(neighborhood.contiguous()), _d(position.contiguous()),
                                _d(out)), "flex_deconv")
    return out


def cuda_group_point_grad(n, grad_out, idx):
    import torch
    B, M, S, C = grad_out.shape
    gp = torch.empty((B, int(n), C), dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_group_point_grad(B, int(n), C, M, S, _d(grad_out.contiguous()), _d(idx.contiguous()), _d(gp)),
         "group_point_grad")
    return gp


def cuda_gather_point_grad(n, out_g, idx):
    import torch
    B, M, _ = out_g.shape
    gp = torch.empty((B, int(n), 3), dtype=torch.float32, device=out_g.device)
    torch.cuda.synchronize()
    _chk(cuda().ref_gather_point_grad(B, int(n), M, _d(out_g.contiguous()), _d(idx.contiguous()), _d(gp)),
         "gather_point_grad")
    return gp
