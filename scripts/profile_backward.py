#!/usr/bin/env python
"""One pass over the backward kernels + NMS at DH3D stage-1 size (8 clouds x 8192 points) for ncu.

    ncu --set full --clock-control none --profile-from-start off -o /tmp/prof_bwd python scripts/profile_backward.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import synth_clouds  # noqa: E402
from dh3d_b200 import tf_ops, user_ops, utils  # noqa: E402


def main():
    torch.cuda.set_device(0)
    B, N, K = 8, 8192, 8
    g = torch.Generator(device="cuda").manual_seed(0)
    pts = synth_clouds(B, N, 0).cuda()
    pos = pts.transpose(1, 2).contiguous()
    ids, _ = user_ops.knn_bruteforce(pos, K)
    nbr = ids.transpose(1, 2).contiguous()
    f = torch.randn((B, 64, N), device="cuda", generator=g)
    th = torch.randn((3, 64, 64), device="cuda", generator=g) / 8
    bi = torch.randn((64, 64), device="cuda", generator=g) / 8
    top = torch.randn((B, 64, N), device="cuda", generator=g)
    _, arg = user_ops.flex_pooling(f, nbr)
    th2 = torch.randn((3, 32), device="cuda", generator=g)
    bi2 = torch.randn((32,), device="cuda", generator=g)
    top32 = torch.randn((B, 32, N), device="cuda", generator=g)
    known = torch.randn((B, 1024, 128), device="cuda", generator=g)
    i3 = torch.randint(0, 1024, (B, N, 3), device="cuda", generator=g, dtype=torch.int32)
    w3 = torch.rand((B, N, 3), device="cuda", generator=g)
    go = torch.randn((B, N, 128), device="cuda", generator=g)
    att = torch.rand((B, N), device="cuda", generator=g)
    dense = (torch.rand((B, N, 3), device="cuda", generator=g) * torch.tensor([12.0, 12.0, 1.5], device="cuda")).contiguous()

    def once():
        user_ops.flex_convolution_grad(f, th, bi, nbr, pos, top)
        user_ops.flex_convolution_transpose(f, pos, nbr, th, bi)
        user_ops.flex_pooling_grad(f, nbr, top, arg)
        user_ops.convolution_pointset_grad(pos, th2, bi2, nbr, top32)
        tf_ops.three_interpolate_grad(known, i3, w3, go)
        utils.batched_nms(dense, att, 0.5, 0.01, 512)

    once()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    once()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
