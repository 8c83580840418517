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
