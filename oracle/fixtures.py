"""The reference's own test fixtures, regenerated without TensorFlow (test infrastructure).

* ``FakePointCloud`` -- user_ops/misc.py:27-84: ``np.random.seed(42)`` at import, then float32
  ``randn`` draws in the order theta, bias, theta_rel, bias_rel, position, features, widened to
  fp64; neighbourhoods = argsort of the fp64 pairwise Euclidean distances, self first.
* ``python_bruteforce`` -- user_ops/test_knn_bruteforce.py:32-40, the only known-answer oracle in
  the reference tree (fp64 distances, argsort / sort, first k).
* ``flexpool_four_point_case`` -- user_ops/test_flex_pooling.py:76-98.
"""
import numpy as np
from scipy.spatial.distance import cdist


def _pairwise(batch_cn):
    pts = np.asarray(batch_cn).T
    return cdist(pts, pts, "euclidean")


class FakePointCloud(object):
    """Regenerates the arrays of user_ops/misc.py:31-66 from a RandomState in the same order."""

    def __init__(self, B, N, K, Din, Dout, Dp, rng=None):
        assert K < N
        rng = np.random.RandomState(42) if rng is None else rng
        self.B, self.N, self.K, self.Din, self.Dout, self.Dp = B, N, K, Din, Dout, Dp

        def draw(shape):
            return rng.randn(*shape).astype(np.float32).astype(np.float64)

        self.theta = draw([Dp, Din, Dout])
        self.bias = draw([Din, Dout])
        self.theta_rel = draw([Din, Dout])
        self.bias_rel = draw([Dout])
        self.position = draw([B, Dp, N])
        self.features = draw([B, Din, N])
        nbrs = [np.argsort(_pairwise(pc), axis=1)[:, :K] for pc in self.position]
        self.neighborhood = np.array(nbrs).transpose(0, 2, 1).astype(np.int32)


def reference_test_cases():
    """The two module-level ``case`` objects of user_ops/test_knn_bruteforce.py:28-29, drawn
    back to back from one seed-42 stream exactly as the test module does at import."""
    rng = np.random.RandomState(42)
    first = FakePointCloud(B=2, N=32, K=4, Din=2, Dout=6, Dp=3, rng=rng)
    second = FakePointCloud(B=1, N=4, K=2, Din=1, Dout=1, Dp=3, rng=rng)
    return first, second


def python_bruteforce(positions, k):
    """positions [B,Dp,N] (any float dtype) -> (ids [B,N,k], dists [B,N,k]) in fp64."""
    ids, dists = [], []
    for pc in positions:
        d = _pairwise(pc)
        ids.append(np.argsort(d, axis=1)[:, :k])
        dists.append(np.sort(d, axis=1)[:, :k])
    return np.array(ids), np.array(dists)


def flexpool_four_point_case():
    """x=[1,2,5,3] on a 4-ring; returns (features [1,1,4] f32, neighborhood [1,4,4] i32).
    Every point's neighbourhood contains index 2 (value 5), so max==5 and argmax==2 everywhere,
    which is what makes the reference's gradient sum to 4 at index 2."""
    x = np.array([[[1], [2], [5], [3]]]).transpose(0, 2, 1).astype(np.float32)
    n = np.array([[[0, 1, 2, 3], [1, 2, 3, 0], [2, 3, 0, 1], [3, 0, 1, 2]]])
    return x, n.transpose(0, 2, 1).astype(np.int32)
