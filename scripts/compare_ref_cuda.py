"""Op-level comparison on the same GPU: this repo's kernels vs the REFERENCE'S OWN CUDA kernels
(oracle/_ref/libdh3d_ref_cuda.so, unmodified sources compiled for sm_100a) -- BASELINE.md 2.2.
Also the FlexConv + kNN sweep of BASELINE.json configs[4] (achieved algorithmic GB/s).

    python scripts/compare_ref_cuda.py [out.json]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dh3d_b200 import ops, tf_ops, user_ops  # noqa: E402
from oracle import ref  # noqa: E402


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main(out_path):
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    have_ref = ref.have_cuda()

    def cloud(B, N):
        return (torch.rand((B, N, 3), device="cuda", generator=g) * 50 - 25).contiguous()

    # ---- k-NN ----
    for (B, N, K) in ((8, 4096, 8), (8, 8192, 8), (8, 8192, 16), (8, 8192, 32), (8, 16384, 8), (8, 32768, 8),
                      (8, 32768, 32), (32, 8192, 8)):
        pts = cloud(B, N)
        pos = pts.transpose(1, 2).contiguous()
        mine = timeit(lambda: user_ops.knn_bruteforce(pos, K))
        r = timeit(lambda: ref.cuda_knn(pos, K), 1, 3) if (have_ref and N <= 8192) else None
        byts = B * (12.0 * N + 8.0 * N * K)
        rows.append({"op": "knn", "B": B, "N": N, "K": K, "ms": mine, "ref_cuda_ms": r,
                     "speedup": (r / mine) if r else None, "algorithmic_GBs": byts / mine / 1e6,
                     "pairs_per_s": B * float(N) * N / mine * 1e3,
                     "note": None if N <= 8192 else "reference kernel refuses N > 8192"})
    # ---- FPS ----
    for (B, N, M) in ((32, 8192, 1024), (8, 8192, 1024), (148, 8192, 1024)):
        pts = cloud(B, N)
        mine = timeit(lambda: tf_ops.farthest_point_sample(M, pts))
        r = timeit(lambda: ref.cuda_fps(M, pts), 1, 3) if have_ref else None
        rows.append({"op": "fps", "B": B, "N": N, "M": M, "ms": mine, "ref_cuda_ms": r,
                     "speedup": (r / mine) if r else None, "rounds_per_s_per_cloud": (M - 1) / mine * 1e3})
    # ---- FlexConv sweep (configs[4]): C=128 ----
    for (B, N, K) in ((8, 4096, 8), (8, 8192, 8), (8, 8192, 16), (8, 8192, 32), (8, 16384, 8), (8, 32768, 8),
                      (4, 32768, 32)):
        C = 128
        pts = cloud(B, N)
        nbr, _ = ops.knn_points(pts, K)
        f = torch.randn((B, N, C), device="cuda", generator=g)
        th = torch.randn((3, C, C), device="cuda", generator=g) / C ** 0.5
        bi = torch.randn((C, C), device="cuda", generator=g) / C ** 0.5
        mine = timeit(lambda: ops.flex_conv(f, th, bi, nbr, pts))
        r = None
        if have_ref:
            fc, pc, nc = f.transpose(1, 2).contiguous(), pts.transpose(1, 2).contiguous(), nbr.transpose(1, 2).contiguous()
            r = timeit(lambda: ref.cuda_flex_conv(fc, pc, nc, th, bi), 1, 3)
        n = B * N
        byts = 4.0 * (n * C + n * C + n * K + 3 * n + 4 * C * C)
        rows.append({"op": "flex_conv", "B": B, "N": N, "K": K, "Cin": C, "Cout": C, "ms": mine,
                     "ref_cuda_ms": r, "speedup": (r / mine) if r else None,
                     "algorithmic_GBs": byts / mine / 1e6, "frac_of_hbm_peak": byts / mine / 1e6 / hbm})
    # ---- DH3D layer shapes ----
    B, N = 32, 8192
    pts = cloud(B, N)
    nbr, _ = ops.knn_points(pts, 8)
    pc, nc = pts.transpose(1, 2).contiguous(), nbr.transpose(1, 2).contiguous()
    for (ci, co) in ((32, 64), (64, 64)):
        f = torch.randn((B, N, ci), device="cuda", generator=g)
        th = torch.randn((3, ci, co), device="cuda", generator=g) / ci ** 0.5
        bi = torch.randn((ci, co), device="cuda", generator=g) / ci ** 0.5
        mine = timeit(lambda: ops.flex_conv(f, th, bi, nbr, pts))
        fc = f.transpose(1, 2).contiguous()
        r = timeit(lambda: ref.cuda_flex_conv(fc, pc, nc, th, bi), 1, 3) if have_ref else None
        n = B * N
        byts = 4.0 * (n * ci + n * co + n * 8 + 3 * n + 4 * ci * co)
        rows.append({"op": "flex_conv", "B": B, "N": N, "K": 8, "Cin": ci, "Cout": co, "ms": mine, "ref_cuda_ms": r,
                     "speedup": (r / mine) if r else None, "algorithmic_GBs": byts / mine / 1e6,
                     "frac_of_hbm_peak": byts / mine / 1e6 / hbm})
    f = torch.randn((B, N, 64), device="cuda", generator=g)
    fc = f.transpose(1, 2).contiguous()
    mine = timeit(lambda: ops.flex_pool(f, nbr))
    r = timeit(lambda: ref.cuda_flex_pool(fc, nc), 1, 3) if have_ref else None
    byts = 4.0 * (2 * B * N * 64 + B * N * 8)
    rows.append({"op": "flex_pool", "B": B, "N": N, "D": 64, "ms": mine, "ref_cuda_ms": r,
                 "speedup": (r / mine) if r else None, "algorithmic_GBs": byts / mine / 1e6,
                 "frac_of_hbm_peak": byts / mine / 1e6 / hbm})
    th2 = torch.randn((3, 32), device="cuda", generator=g)
    bi2 = torch.randn((32,), device="cuda", generator=g)
    mine = timeit(lambda: ops.conv_pointset(pts, th2, bi2, nbr))
    r = timeit(lambda: ref.cuda_conv_pointset(pc, nc, th2, bi2), 1, 3) if have_ref else None
    rows.append({"op": "conv_pointset", "B": B, "N": N, "ms": mine, "ref_cuda_ms": r,
                 "speedup": (r / mine) if r else None})
    kp = tf_ops.farthest_point_sample(1024, pts).unsqueeze(2).contiguous()
    f128 = torch.randn((B, N, 128), device="cuda", generator=g)
    mine = timeit(lambda: tf_ops.group_point(f128, kp))
    r = timeit(lambda: ref.cuda_group_point(f128, kp), 1, 3) if have_ref else None
    rows.append({"op": "group_point", "B": B, "N": N, "M": 1024, "C": 128, "ms": mine, "ref_cuda_ms": r,
                 "speedup": (r / mine) if r else None})
    # ---- backward passes / FlexDeconv / NMS, DH3D layer shapes at 8 clouds ----
    from dh3d_b200 import utils as dutils
    Bb = 8
    ptsb, nbrb = pts[:Bb].contiguous(), nbr[:Bb].contiguous()
    pcb, ncb = ptsb.transpose(1, 2).contiguous(), nbrb.transpose(1, 2).contiguous()
    for (ci, co) in ((32, 64), (64, 64)):
        fcb = torch.randn((Bb, ci, N), device="cuda", generator=g)
        th = torch.randn((3, ci, co), device="cuda", generator=g) / ci ** 0.5
        bi = torch.randn((ci, co), device="cuda", generator=g) / ci ** 0.5
        top = torch.randn((Bb, co, N), device="cuda", generator=g)
        mine = timeit(lambda: user_ops.flex_convolution_grad(fcb, th, bi, ncb, pcb, top))
        r = timeit(lambda: ref.cuda_flex_conv_grad(fcb, th, bi, ncb, pcb, top), 1, 2) if have_ref else None
        n = Bb * N
        byts = 4.0 * (2 * n * ci + n * co + n * 8 + 3 * n + 2 * 4 * ci * co)   # f, top, nbr, xyz in; grad_f out; params
        rows.append({"op": "flex_conv_grad", "B": Bb, "N": N, "K": 8, "Cin": ci, "Cout": co, "ms": mine,
                     "ref_cuda_ms": r, "speedup": (r / mine) if r else None, "algorithmic_GBs": byts / mine / 1e6,
                     "frac_of_hbm_peak": byts / mine / 1e6 / hbm,
                     "note": "reference-layout entry (4 layout transposes inside the timed call)"})
        mine = timeit(lambda: user_ops.flex_convolution_transpose(fcb, pcb, ncb, th, bi))
        r = timeit(lambda: ref.cuda_flex_deconv(fcb, pcb, ncb, th, bi), 1, 2) if have_ref else None
        rows.append({"op": "flex_deconv", "B": Bb, "N": N, "K": 8, "Cin": ci, "Cout": co, "ms": mine,
                     "ref_cuda_ms": r, "speedup": (r / mine) if r else None})
    fcb = torch.randn((Bb, 64, N), device="cuda", generator=g)
    top = torch.randn((Bb, 64, N), device="cuda", generator=g)
    _, arg = user_ops.flex_pooling(fcb, ncb)
    mine = timeit(lambda: user_ops.flex_pooling_grad(fcb, ncb, top, arg))
    r = timeit(lambda: ref.cuda_flex_pool_grad(fcb, ncb, top, arg), 1, 3) if have_ref else None
    rows.append({"op": "flex_pool_grad", "B": Bb, "N": N, "D": 64, "ms": mine, "ref_cuda_ms": r,
                 "speedup": (r / mine) if r else None, "algorithmic_GBs": 4.0 * 3 * Bb * N * 64 / mine / 1e6})
    th2c, top32 = th2.contiguous(), torch.randn((Bb, 32, N), device="cuda", generator=g)
    mine = timeit(lambda: user_ops.convolution_pointset_grad(pcb, th2c, bi2, ncb, top32))
    r = timeit(lambda: ref.cuda_conv_pointset_grad(pcb, th2c, bi2, ncb, top32), 1, 2) if have_ref else None
    rows.append({"op": "conv_pointset_grad", "B": Bb, "N": N, "ms": mine, "ref_cuda_ms": r,
                 "speedup": (r / mine) if r else None})
    go = torch.randn((B, 1024, 1, 128), device="cuda", generator=g)
    mine = timeit(lambda: tf_ops.group_point_grad(f128, kp, go))
    r = timeit(lambda: ref.cuda_group_point_grad(N, go, kp), 1, 3) if have_ref else None
    rows.append({"op": "group_point_grad", "B": B, "N": N, "M": 1024, "C": 128, "ms": mine, "ref_cuda_ms": r,
                 "speedup": (r / mine) if r else None})
    known = torch.randn((B, 1024, 128), device="cuda", generator=g)
    d3, i3 = tf_ops.three_nn(pts, tf_ops.gather_point(pts, kp.squeeze(2).contiguous()))
    w3 = torch.softmax(-d3, dim=2).contiguous()
    mine = timeit(lambda: tf_ops.three_interpolate_grad(known, i3, w3, f128))
    rows.append({"op": "three_interpolate_grad", "B": B, "N": N, "M": 1024, "C": 128, "ms": mine,
                 "ref_cuda_ms": None, "speedup": None, "algorithmic_GBs": 4.0 * B * (N * 128 + 1024 * 128 + 6 * N) / mine / 1e6,
                 "note": "CPU-only op in the reference"})
    att = torch.rand((B, N), device="cuda", generator=g)
    dense = (torch.rand((B, N, 3), device="cuda", generator=g) * torch.tensor([12.0, 12.0, 1.5], device="cuda")).contiguous()
    mine = timeit(lambda: dutils.batched_nms(dense, att, 0.5, 0.01, 512))
    rows.append({"op": "keypoint_nms", "B": B, "N": N, "ms": mine, "ref_cuda_ms": None, "speedup": None,
                 "clouds_per_s": B / mine * 1e3, "note": "host numpy + sklearn ball tree in the reference (core/utils.py:15-43)"})
    res = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": hbm, "rows": rows,
           "note": "ref_cuda_ms = the reference's unmodified CUDA kernels compiled for sm_100a (oracle/_ref), "
                   "same inputs, same GPU, CUDA events, median of 3-5; three_nn/three_interpolate have no "
                   "reference GPU kernel (CPU-only ops in the reference)."}
    with open(out_path, "w") as f_:
        json.dump(res, f_, indent=1)
    for r_ in rows:
        print(json.dumps(r_))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_cuda_compare.json"))
