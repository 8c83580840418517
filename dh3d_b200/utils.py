"""Mirror of the one hot-path helper of the reference's ``core/utils.py``: keypoint NMS.

    single_nms(xyz, attention, nms_radius, min_response_ratio, max_keypoints, remove_noise=True)
        -> (num_keypoints, max_indices)                                    core/utils.py:15-43

Same name, argument order and return tuple over CUDA tensors (the reference: numpy + an sklearn ball
tree on the host, one cloud per call from evaluate/local_eval/localdesc_extract.py:92-98).  Unlike the
reference, ``attention`` is not modified in place.  ``batched_nms`` is the same computation for a whole
batch in one launch sequence (no host round trip per cloud).
"""
import ctypes

import torch

from ._lib import call, check, query, stream_ptr, workspace

f32, i32 = torch.float32, torch.int32


def batched_nms(xyz, attention, nms_radius, min_response_ratio, max_keypoints, remove_noise=True):
    """xyz [B,N,3], attention [B,N] -> (indices [B,max_keypoints] i32 padded with -1, counts [B] i32)."""
    check(xyz, f32, "xyz", 3)
    check(attention, f32, "attention", 2)
    B, N, _ = xyz.shape
    out = torch.empty((B, int(max_keypoints)), dtype=i32, device=xyz.device)
    cnt = torch.empty((B,), dtype=i32, device=xyz.device)
    ws, wp, wn = workspace(query("dh3d_keypoint_nms_workspace_bytes", B, N), xyz.device)
    call("dh3d_keypoint_nms", check(xyz, f32, "xyz"), check(attention, f32, "attention"), B, N,
         ctypes.c_float(nms_radius), ctypes.c_float(min_response_ratio), int(max_keypoints),
         int(bool(remove_noise)), check(out, i32, "out"), check(cnt, i32, "cnt"), wp, wn,
         stream_ptr(xyz.device))
    return out, cnt


def single_nms(xyz, attention, nms_radius, min_response_ratio, max_keypoints, remove_noise=True):
    """xyz [N,3], attention [N] -> (num_keypoints int, max_indices i32 tensor [num_keypoints])."""
    out, cnt = batched_nms(xyz.reshape(1, -1, 3).contiguous(), attention.reshape(1, -1).contiguous(),
                           nms_radius, min_response_ratio, max_keypoints, remove_noise)
    n = int(cnt.item())
    return n, out[0, :n]
