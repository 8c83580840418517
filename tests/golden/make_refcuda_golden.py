"""Generates tests/golden/refcuda_*.npz: outputs of the REFERENCE'S OWN CUDA kernels (unmodified
sources compiled for sm_100a by oracle/build_ref.py) on small seeded inputs.  Must run on a GPU:

    gpurun -- 'python tests/golden/make_refcuda_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/.  tests/test_golden.py replays them against
the CPU oracle without a GPU, so the oracle stays pinned to the real reference in CPU-only runs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lattice_cloud, make_cloud  # noqa: E402
from oracle import ref  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.RandomState(20260925)
    # kNN: random, lattice (ties) and duplicated clouds at several (T,V) dispatch sizes
    for N, K in ((40, 4), (200, 8), (600, 8), (1500, 8), (3000, 16)):
        pts = make_cloud(rng, 3, N)
        pts[1] = lattice_cloud(rng, 1, N)[0]
        pts[2, N - N // 3:] = pts[2, :N // 3]
        pos = np.ascontiguousarray(pts.transpose(0, 2, 1))
        ids, d = ref.cuda_knn(cu(pos), K)
        np.savez_compressed(os.path.join(out_dir, "refcuda_knn_n%d_k%d.npz" % (N, K)), positions=pos,
                            ids=ids.cpu().numpy(), dists=d.cpu().numpy())
    # FPS
    for N, M in ((700, 64), (2048, 256), (3000, 100)):
        pts = make_cloud(rng, 3, N)
        pts[1] = lattice_cloud(rng, 1, N, step=1.0, side=5)[0]
        pts[2, N // 2:] = pts[2, :N - N // 2]
        idx = ref.cuda_fps(M, cu(pts))
        np.savez_compressed(os.path.join(out_dir, "refcuda_fps_n%d_m%d.npz" % (N, M)), xyz=pts,
                            idx=idx.cpu().numpy())
    # ball query (incl. queries with no hit -> leaked nearest)
    xyz1, xyz2 = make_cloud(rng, 2, 500, extent=2.0), make_cloud(rng, 2, 700, extent=2.5)
    idx, cnt = ref.cuda_query_ball_point(0.25, 8, cu(xyz1), cu(xyz2))
    np.savez_compressed(os.path.join(out_dir, "refcuda_ball.npz"), xyz1=xyz1, xyz2=xyz2, radius=0.25, nsample=8,
                        idx=idx.cpu().numpy(), cnt=cnt.cpu().numpy())
    # flex ops on a small layer
    B, N, K, Din, Dout = 2, 256, 8, 16, 24
    pts = make_cloud(rng, B, N, extent=5.0)
    pos = np.ascontiguousarray(pts.transpose(0, 2, 1))
    nb = ref.cuda_knn(cu(pos), K)[0].cpu().numpy().transpose(0, 2, 1).copy()
    f = rng.randn(B, Din, N).astype(np.float32)
    th, bi = (rng.randn(3, Din, Dout) / 4).astype(np.float32), (rng.randn(Din, Dout) / 4).astype(np.float32)
    th2, bi2 = rng.randn(Din, Dout).astype(np.float32), rng.randn(Dout).astype(np.float32)
    fc = ref.cuda_flex_conv(cu(f), cu(pos), cu(nb), cu(th), cu(bi)).cpu().numpy()
    po, pa = ref.cuda_flex_pool(cu(f), cu(nb))
    cp = ref.cuda_conv_pointset(cu(f), cu(nb), cu(th2), cu(bi2)).cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "refcuda_flex.npz"), features=f, position=pos, neighborhood=nb,
                        theta=th, bias=bi, theta_rel=th2, bias_rel=bi2, flex_conv=fc, pool=po.cpu().numpy(),
                        argmax=pa.cpu().numpy(), conv_pointset=cp)
    print("wrote", sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
