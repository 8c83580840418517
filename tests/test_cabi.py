"""CPU tests of the drop-in boundary: the shared library loads without a GPU, exports exactly the
symbols include/dh3d_b200.h declares, validates arguments before touching CUDA, and the Python
boundary refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from dh3d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "dh3d_b200.h")).read()
    return sorted(set(re.findall(r"DH3D_API\s+[\w\s\*]+?\b(dh3d_\w+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libdh3d_b200.so does not export %s" % n
    assert sorted(_lib.exported_symbols()) == names, "python binding table out of sync with the header"
    assert lib.dh3d_version() == 100


def test_error_strings():
    assert _lib.error_string(0) == "ok"
    assert "NULL" in _lib.error_string(-1)
    assert "workspace" in _lib.error_string(-4)


def test_argument_validation_happens_before_any_cuda_call():
    lib = _lib.lib()
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert lib.dh3d_knn_bruteforce(null, 1, 3, 8, 2, one, one, one, 1 << 20, null) == -1
    assert lib.dh3d_knn_bruteforce(one, 1, 3, 0, 2, one, one, one, 1 << 20, null) == -2
    assert lib.dh3d_knn_bruteforce(one, 1, 2, 8, 2, one, one, one, 1 << 20, null) == -3   # Dp != 3
    assert lib.dh3d_knn_bruteforce(one, 1, 3, 8, 65, one, one, one, 1 << 20, null) == -3  # K > 64
    assert lib.dh3d_knn_bruteforce(one, 1, 3, 8, 2, one, one, one, 16, null) == -4        # workspace
    assert lib.dh3d_farthest_point_sample(0, 8, 2, one, one, null) == -2
    assert lib.dh3d_farthest_point_sample(1, 8, 0, one, one, null) == 0                    # m == 0: no-op
    assert lib.dh3d_flex_conv_pm(one, one, one, one, one, one, 1, 8, 4, 6, 8, null, null, null, 0,
                                 one, 1 << 30, null) == -3                                  # Din % 4
    assert lib.dh3d_conv_pointset(one, one, one, one, one, 1, 8, 4, 65, 8, null) == -3      # Din > 64
    assert lib.dh3d_linear(one, 6, one, null, null, 0, one, 8, 4, 6, 8, null) == -2         # K % 4
    assert lib.dh3d_three_nn(1, 0, 4, one, one, one, one, null) == -2
    # backward passes / deconv / NMS
    assert lib.dh3d_flex_conv_grad(one, one, one, one, one, one, one, one, null, 1, 8, 4, 6, 8, one, 1 << 30, null) == -1
    assert lib.dh3d_flex_conv_grad(one, one, one, one, one, one, one, one, one, 1, 8, 4, 6, 8, one, 16, null) == -4
    assert lib.dh3d_flex_conv_grad_pm(one, one, one, one, one, one, one, one, one, 1, 8, 4, 6, 8, one, 1 << 30, null) == -3
    assert lib.dh3d_flex_pool_grad(one, one, one, 1, 0, 4, null) == -2
    assert lib.dh3d_conv_pointset_grad(one, one, one, one, one, one, one, 1, 8, 4, 3, 32, one, 16, null) == -4
    assert lib.dh3d_flex_deconv(one, one, one, one, one, one, 0, 8, 4, 6, 8, one, 1 << 30, null) == -2
    assert lib.dh3d_group_point_grad(1, 8, 4, 0, 1, one, one, one, null) == -2
    assert lib.dh3d_three_interpolate_grad(1, 8, 4, 2, one, one, null, one, null) == -1
    assert lib.dh3d_keypoint_nms(one, one, 1, 40, 0.5, 0.01, 16, 1, one, one, one, 1 << 30, null) == -3   # N < 50
    assert lib.dh3d_keypoint_nms(one, one, 1, 400, 0.5, 0.01, 16, 1, one, one, one, 16, null) == -4
    # fused inference entry points: squeeze/excite, prepacked FlexConv, two-branch join
    f = ctypes.c_float
    assert lib.dh3d_se_pool_excite(null, one, one, one, one, one, one, 1, 8, 4, 64, 16, null) == -1
    assert lib.dh3d_se_pool_excite(one, one, one, one, one, one, one, 1, 0, 4, 64, 16, null) == -2
    assert lib.dh3d_se_pool_excite(one, one, one, one, one, one, one, 1, 8, 4, 96, 24, null) == -3      # C not 64/128
    assert lib.dh3d_se_pool_excite(one, one, one, one, one, one, one, 1, 8, 4, 64, 32, null) == -3      # H != C/4
    assert lib.dh3d_flex_conv_prepack_bytes(64, 128) >= 2 * 4 * 64 * 128 * 4 + 128 * 4
    assert lib.dh3d_flex_conv_prepack(null, one, null, null, null, 64, 128, one, null) == -1
    assert lib.dh3d_flex_conv_prepack(one, one, null, null, null, 6, 128, ctypes.c_void_p(256), null) == -3
    assert lib.dh3d_flex_conv_prepack(one, one, null, null, null, 64, 128, ctypes.c_void_p(16), null) == -5  # alignment
    assert lib.dh3d_flex_conv_pm_packed(one, null, one, one, one, 1, 8, 4, 64, 64, null, 0, one, 1 << 30, null) == -1
    assert lib.dh3d_flex_conv_pm_packed(one, ctypes.c_void_p(256), one, one, one, 1, 8, 65, 64, 64, null, 0, one,
                                        1 << 30, null) == -3                                             # K > 64
    assert lib.dh3d_flex_conv_pm_packed(one, ctypes.c_void_p(256), one, one, one, 1, 8, 4, 64, 64, null, 0, one, 0,
                                        null) == -4
    assert lib.dh3d_linear_join_packed(one, 192, one, null, null, 1, one, 64, one, null, null, 1, one, 64, null, 64,
                                       f(1e-8), 8, 192, 64, 64, null) == -3                             # N != 128
    assert lib.dh3d_linear_join_packed(one, 192, one, null, null, 1, null, 64, one, null, null, 1, one, 128, null,
                                       128, f(1e-8), 8, 192, 64, 128, null) == -1
    assert lib.dh3d_linear_join_packed(one, 190, one, null, null, 1, one, 64, one, null, null, 1, one, 128, null,
                                       128, f(1e-8), 8, 190, 64, 128, null) == -2                       # K % 4
    assert lib.dh3d_three_nn_ws_presorted(1, 64, 8, null, one, one, one, one, 1 << 30, null) == -1
    assert lib.dh3d_three_nn_ws_presorted(1, 64, 8, ctypes.c_void_p(256), one, one, one, ctypes.c_void_p(256), 16,
                                          null) == -4                                                   # workspace
    assert lib.dh3d_three_nn_presorted2(1, 64, 8, ctypes.c_void_p(256), null, one, one, null) == -1
    assert lib.dh3d_farthest_point_sample_presorted(1, 64, 8, null, one, null) == -1
    assert lib.dh3d_farthest_point_sample_presorted(1, 9000, 8, one, one, null) == -3                   # n > 8192
    assert lib.dh3d_knn_sort_pm(one, 1, 64, ctypes.c_void_p(256), 16, null) == -4
    assert lib.dh3d_knn_query_sorted(null, 1, 64, 8, one, one, null) == -1
    # chained two-layer entry: x, ldx, packed1, scale1, shift1, act1, packed2, scale2, shift2, act2, y, ldy, M, K1, N1, N2
    assert lib.dh3d_linear_chain_packed(one, 128, one, null, null, 1, one, null, null, 1, one, 256, 8, 128, 64, 256,
                                        null) == -3                                                     # N1 != 128
    assert lib.dh3d_linear_chain_packed(one, 256, one, null, null, 1, one, null, null, 1, one, 256, 8, 256, 128, 256,
                                        null) == -3                                                     # K1 > 128
    assert lib.dh3d_linear_chain_packed(one, 128, one, null, null, 1, one, null, null, 1, one, 512, 8, 128, 128, 512,
                                        null) == -3                                                     # N2 > 256
    assert lib.dh3d_linear_chain_packed(null, 128, one, null, null, 1, one, null, null, 1, one, 256, 8, 128, 128, 256,
                                        null) == -1
    assert lib.dh3d_linear_chain_packed(one, 128, one, null, null, 1, null, null, null, 1, one, 256, 8, 128, 128, 256,
                                        null) == -1
    assert lib.dh3d_linear_chain_packed(one, 128, one, null, null, 1, one, null, null, 1, one, 200, 8, 128, 128, 256,
                                        null) == -2                                                     # ldy < N2
    assert lib.dh3d_netvlad_workspace_bytes(2, 100, 128, 64, 256) == 0                      # unsupported dims
    assert lib.dh3d_netvlad_workspace_bytes(2, 100, 256, 64, 256) > 0


def test_workspace_queries():
    lib = _lib.lib()
    # sorted float4 points + one (min,max) float4 pair per 32-point chunk + one pair per 16 chunks
    assert lib.dh3d_knn_workspace_bytes(2, 8192) == 2 * (8192 * 16 + 256 * 32 + 16 * 32)
    assert lib.dh3d_knn_workspace_bytes(1, 4) == 32 * 16 + 1 * 32 + 1 * 32      # one padded chunk, one super-chunk
    assert lib.dh3d_knn_workspace_bytes(1, 9000) == 9024 * 16 + 282 * 32 + 18 * 32
    assert lib.dh3d_flex_conv_pm_workspace_bytes(1, 128, 8, 32, 64) >= 128 * 4 * 32 * 4 + 4 * 32 * 64 * 4
    assert lib.dh3d_flex_conv_workspace_bytes(1, 32, 4, 2, 6) > 0     # padded odd dims are accepted


def test_python_boundary_refuses_cpu_tensors():
    from dh3d_b200 import ops, user_ops
    with pytest.raises(_lib.Dh3dError, match="CUDA tensor"):
        user_ops.knn_bruteforce(torch.zeros(1, 3, 8), 2)
    with pytest.raises(_lib.Dh3dError, match="CUDA tensor"):
        ops.farthest_point_sample(4, torch.zeros(1, 8, 3))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdh3d_b200.so")
    with pytest.raises(_lib.Dh3dError, match="no CPU fallback"):
        _lib.lib()


def test_model_parameter_names_follow_reference_scopes():
    from dh3d_b200.configs import full_config
    from dh3d_b200.model import DH3D
    names = dict(DH3D(full_config()).named_parameters())
    for n, shape in {
        "local.initconv.position_theta": (3, 32),
        "local.stage1.flexconv_0.position_theta": (3, 32, 64),
        "local.stage1.flexconv_1.position_bias": (64, 64),
        "local.stage1.se.f1.tfconv0.W": (1, 1, 64, 16),
        "local.before_stage2_conv1d.tfconv0.bn.gamma": (64,),
        "local.stage2.flexconv_1.feature_bias": (128, 1),
        "local.stage2.concat_conv1d.tfconv0.W": (1, 1, 192, 128),
        "local.local_stage1_shortcut.tfconv0.W": (1, 1, 64, 128),
        "detection_block_reliable.detec_conv2.W": (1, 1, 256, 1024),
        "detection_block_reliable.detec_conv_fc.W": (1, 1, 1024, 1),
        "global_before_assemble.flexconv_0.position_theta": (3, 128, 256),
        "globalatt.detec_conv0.W": (1, 1, 256, 1024),
        "netvlad.cluster_weights": (256, 64),
        "netvlad.cluster_weights2": (1, 256, 64),
        "netvlad.hidden1_weights": (16384, 256),
        "netvlad.gating_bn.variance_ema": (256,),
    }.items():
        assert tuple(names[n].shape) == shape, n
