"""Place-recognition retrieval after the descriptor all-gather (reference
evaluate/global_eval/evaluation_retrieval.py:29-58,129-169): k nearest reference descriptors per
query (Euclidean, 256-D) and recall@N / top-1% against the 25 m UTM ground truth.

``retrieve_topk`` runs on the GPU through the C ABI (Gram matrix with the GEMM kernel, a warp-per-query
selection of k + 8 candidates, then a re-rank of those by their exactly computed distances -- the Gram form alone
cannot order near-identical descriptors); the recall bookkeeping is host numpy like the reference."""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from ._lib import call, check, stream_ptr


def retrieve_topk(ref_desc, query_desc, k):
    """ref_desc [R,D], query_desc [Q,D] CUDA fp32 -> (idx [Q,k] i32, dist2 [Q,k] f32), ascending."""
    R, D = ref_desc.shape
    Q = query_desc.shape[0]
    k = min(int(k), R)
    Rp = (R + 3) // 4 * 4  # the GEMM wants a column count that is a multiple of 4: zero-pad ref^T
    ref_t = torch.zeros((D, Rp), dtype=torch.float32, device=ref_desc.device)
    ref_t[:, :R] = ops.transpose_pm_to_cm(ref_desc.reshape(1, R, D).contiguous()).reshape(D, R)
    gram = ops.linear(query_desc.contiguous(), ref_t)     # [Q, Rp]
    qn = (query_desc * query_desc).sum(1).contiguous()
    rn = (ref_desc * ref_desc).sum(1).contiguous()
    idx = torch.empty((Q, k), dtype=torch.int32, device=ref_desc.device)
    val = torch.empty((Q, k), dtype=torch.float32, device=ref_desc.device)
    if k <= 24:
        cand = torch.empty((Q, 32), dtype=torch.int32, device=ref_desc.device)
        cval = torch.empty((Q, 32), dtype=torch.float32, device=ref_desc.device)
        call("dh3d_topk_l2_exact", check(gram, torch.float32, "gram"), Rp, check(qn, torch.float32, "qn"),
             check(rn, torch.float32, "rn"), check(query_desc.contiguous(), torch.float32, "query"),
             check(ref_desc.contiguous(), torch.float32, "ref"), Q, R, D, k, check(idx, torch.int32, "idx"),
             check(val, torch.float32, "val"), check(cand, torch.int32, "cand"), check(cval, torch.float32, "cand_val"),
             stream_ptr(ref_desc.device))
        return idx, val
    call("dh3d_topk_l2", check(gram, torch.float32, "gram"), Rp, check(qn, torch.float32, "qn"),
         check(rn, torch.float32, "rn"), Q, R, k, check(idx, torch.int32, "idx"),
         check(val, torch.float32, "val"), stream_ptr(ref_desc.device))
    return idx, val


def is_gt_match_2d(queries, ref, distance_thresh=25.0):
    q = np.stack([np.asarray(queries["northing"]), np.asarray(queries["easting"])], 0)
    r = np.stack([np.asarray(ref["northing"]), np.asarray(ref["easting"])], 0)
    return np.linalg.norm(q[:, :, None] - r[:, None, :], axis=0) < distance_thresh


def recall_from_indices(indices, gt_matches, num_ref):
    """evaluation_retrieval.py:43-53: (recall@1..k over valid queries, top-1% rate)."""
    threshold = max(int(round(num_ref / 100.0)), 1)
    tp = gt_matches[np.arange(len(indices))[:, None], indices]
    valid = np.any(gt_matches, axis=1)
    recall = np.mean(np.cumsum(tp, axis=1)[valid] > 0, axis=0)
    one_percent = np.mean(np.any(tp[:, 0:threshold], axis=1)[valid])
    return recall, one_percent, int(valid.sum())
