"""Host-side readers for the reference's on-disk formats: Oxford `.bin` clouds,
the `*.pickle` sequence dictionaries, and the fixed-size cloud preparation of the evaluation
data flows.  numpy/scipy only (the reference uses open3d for the outlier filter)."""
import pickle

import numpy as np
from scipy.spatial import cKDTree


def load_single_pcfile(filename, dim=3, dtype=np.float32):
    """core/utils.py:145-148: raw little-endian [N,dim] -> xyz [N,3]."""
    pc = np.fromfile(filename, dtype=dtype)
    return pc.reshape(pc.shape[0] // dim, dim)[:, 0:3]


def get_sets_dict(filename):
    """core/utils.py:46-50: {sequence: [{'query': 'seq/id', 'northing': .., 'easting': ..}, ...]}."""
    with open(filename, "rb") as handle:
        return pickle.load(handle)


def remove_noise(pcd, nb_points=4, radius=1.0):
    """core/utils.py:173-177 (open3d remove_radius_outlier): keep points with more than
    ``nb_points`` points (itself included) within ``radius``.  Returns the kept indices."""
    tree = cKDTree(pcd)
    counts = tree.query_ball_point(pcd, r=radius, return_length=True)
    return np.nonzero(counts > nb_points)[0]


def get_fixednum_pcd(cloud, targetnum, rng=None, randsample=True, sortby_dis=True):
    """core/utils.py:87-110: outlier removal, then crop to the ``targetnum`` points nearest the
    centroid and shuffle, or pad short clouds with DUPLICATED points (randsample) / 1e5 points."""
    rng = np.random if rng is None else rng
    cloud = cloud[remove_noise(cloud), :]
    ori_num = cloud.shape[0]
    if cloud.shape[0] > targetnum:
        if sortby_dis:
            centroid = np.mean(cloud, axis=0)
            dis = np.sum(np.square(cloud - centroid), axis=1)
            cloud = cloud[np.argsort(dis)[0:targetnum], :3]
        cloud = cloud[rng.choice(cloud.shape[0], targetnum, replace=False), :]
        ori_num = targetnum
    else:
        num_to_pad = targetnum - cloud.shape[0]
        if randsample:
            pad = cloud[rng.choice(cloud.shape[0], size=num_to_pad, replace=True), :]
        else:
            pad = np.ones([num_to_pad, 3], dtype=np.float32) * 100000
        cloud = np.concatenate((cloud, pad), axis=0)
    return np.ascontiguousarray(cloud, dtype=np.float32), ori_num


def synth_lidar_clouds(batch, n_points=8192, seed=0, pad_fraction=0.15):
    """Seeded LiDAR-like synthetic clouds (SURVEY 8d's second input set; benchmarks have no dataset access):
    a vehicle-centred scan with the statistics that stress the spatial kernels and that U(-25,25)^3 lacks --

      * a ground plane (55 % of the points) whose areal density falls off as 1/r^2 (radial pdf ~ 1/r, 2..30 m),
      * vertical structures (30 %): ~14 planar facades / poles at 4..28 m, 2..8 m wide, 2..7 m high,
      * vegetation / clutter blobs (15 %): Gaussian clusters, sigma 0.4..1.2 m,
      * z extent ~10 m against +-30 m in x/y (a flat slab, not a cube),
      * the last ``pad_fraction`` of every cloud are DUPLICATES of earlier points -- what get_fixednum_pcd
        (core/utils.py:103-106) appends to short clouds: exact k-NN / FPS ties.

    Returns float32 [batch, n_points, 3]."""
    out = np.empty((batch, n_points, 3), np.float32)
    for b in range(batch):
        rng = np.random.RandomState(977 * seed + b + 1)
        n_pad = int(round(n_points * pad_fraction))
        n_real = n_points - n_pad
        n_ground = int(0.55 * n_real)
        n_struct = int(0.30 * n_real)
        n_veg = n_real - n_ground - n_struct
        r = 2.0 * (30.0 / 2.0) ** rng.rand(n_ground)              # pdf ~ 1/r on [2, 30]
        a = rng.rand(n_ground) * 2 * np.pi
        ground = np.stack([r * np.cos(a), r * np.sin(a), -1.7 + 0.03 * rng.randn(n_ground)], 1)
        parts = [ground]
        n_obj = 14
        counts = rng.multinomial(n_struct, np.ones(n_obj) / n_obj)
        for cnt in counts:
            rr, aa = 4.0 + 24.0 * rng.rand(), rng.rand() * 2 * np.pi
            centre = np.array([rr * np.cos(aa), rr * np.sin(aa)])
            yaw = rng.rand() * np.pi
            w, h = 0.2 + 7.8 * rng.rand() ** 2, 2.0 + 5.0 * rng.rand()
            u = (rng.rand(cnt) - 0.5) * w
            pts = np.stack([centre[0] + u * np.cos(yaw), centre[1] + u * np.sin(yaw), -1.7 + h * rng.rand(cnt)], 1)
            parts.append(pts + 0.02 * rng.randn(cnt, 3))
        n_blob = 10
        counts = rng.multinomial(n_veg, np.ones(n_blob) / n_blob)
        for cnt in counts:
            rr, aa = 3.0 + 25.0 * rng.rand(), rng.rand() * 2 * np.pi
            c = np.array([rr * np.cos(aa), rr * np.sin(aa), -0.5 + 2.5 * rng.rand()])
            parts.append(c + (0.4 + 0.8 * rng.rand()) * rng.randn(cnt, 3))
        real = np.concatenate(parts, 0)[:n_real]
        real = real[rng.permutation(real.shape[0])]
        pad = real[rng.choice(real.shape[0], size=n_pad, replace=True)]
        out[b] = np.concatenate([real, pad], 0).astype(np.float32)
    return out
