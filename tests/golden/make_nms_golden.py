"""Generates tests/golden/nms_*.npz by running the REFERENCE'S OWN `single_nms` (core/utils.py:15-43).

The module `core/utils.py` imports tensorpack / open3d / termcolor at the top, which are absent here, so the
function's source is extracted from the file with `ast` and executed unmodified against numpy + scikit-learn
(both present in this container).  Run from the repo root in the build container:

    python tests/golden/make_nms_golden.py [/root/reference]
"""
import ast
import os
import sys

import numpy as np


def load_reference_single_nms(root):
    path = os.path.join(root, "core", "utils.py")
    src = open(path).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "single_nms")
    code = "\n".join(src.split("\n")[fn.lineno - 1:fn.end_lineno])
    ns = {"np": np}
    exec(compile(code, path, "exec"), ns)
    return ns["single_nms"]


def make_cloud(seed, n_slab, n_out):
    rng = np.random.RandomState(seed)
    slab = rng.uniform([0, 0, 0], [12, 12, 1.5], (n_slab, 3))
    outliers = rng.uniform(-30, 30, (n_out, 3))                 # isolated points: the remove_noise branch
    xyz = np.concatenate([slab, outliers]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    att = rng.rand(len(xyz)).astype(np.float32)
    return xyz, att


CASES = [  # name, seed, n_slab, n_out, nms_radius, min_response_ratio, max_keypoints, remove_noise
    ("a", 1, 4000, 96, 0.5, 0.01, 512, True),      # the reference's local-eval defaults (localdesc_extract.py:169-171)
    ("b", 2, 4000, 96, 1.0, 0.3, 4096, True),       # > 50 points inside the radius: the 50-NN truncation matters
    ("c", 3, 2000, 48, 0.5, 0.01, 4096, False),
]

if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    single_nms = load_reference_single_nms(root)
    here = os.path.dirname(os.path.abspath(__file__))
    for name, seed, ns, no, rad, ratio, kp, noise in CASES:
        xyz, att = make_cloud(seed, ns, no)
        num, idx = single_nms(xyz, att.copy(), rad, ratio, kp, remove_noise=noise)
        np.savez_compressed(os.path.join(here, "nms_%s.npz" % name), xyz=xyz, attention=att,
                            params=np.array([rad, ratio, kp, int(noise)], np.float64),
                            num_keypoints=np.int32(num), max_indices=np.asarray(idx, np.int32))
        print(name, "N", len(xyz), "keypoints", num)
