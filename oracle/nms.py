"""numpy restatement of the reference's keypoint NMS -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows core/utils.py:15-43 `single_nms` line by line, with the sklearn ball tree (:17-18) replaced by an
exhaustive fp64 distance matrix + stable argsort (sklearn works in float64 and returns neighbours in ascending
distance; the order among exactly equal distances is unspecified there and index order here).  Pinned against
the reference function itself: tests/golden/nms_*.npz are outputs of the unmodified `single_nms` source run
with scikit-learn by tests/golden/make_nms_golden.py.
"""
import numpy as np


def single_nms(xyz, attention, nms_radius, min_response_ratio, max_keypoints, remove_noise=True):
    xyz64 = np.asarray(xyz, np.float64)
    attention = np.array(attention, np.float32, copy=True)          # the reference mutates its argument (:22)
    d2 = ((xyz64[:, None, :] - xyz64[None, :, :]) ** 2).sum(-1)
    indices = np.argsort(d2, axis=1, kind="stable")[:, :50]          # :17-18, n_neighbors=50
    distances = np.sqrt(np.take_along_axis(d2, indices, axis=1))
    if remove_noise:                                                 # :19-22
        attention[distances[:, 7] > 2.0] = 0.0
    knn_attention = attention[indices]                               # :24
    knn_attention[distances > nms_radius] = 0.0                      # :25-26
    is_max = np.where(np.argmax(knn_attention, axis=1) == 0)[0]      # :27
    attention_thresh = np.max(attention) * min_response_ratio       # :30
    is_max_attention = [(attention[m], m) for m in is_max if attention[m] > attention_thresh]   # :32
    is_max_attention = sorted(is_max_attention, reverse=True)        # :33
    max_indices = [int(m[1]) for m in is_max_attention][:max_keypoints]   # :35-40
    return len(max_indices), np.asarray(max_indices, np.int32)
