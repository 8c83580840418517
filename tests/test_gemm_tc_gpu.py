"""GPU tests of the tcgen05 (3xTF32) GEMM path against fp64 and against the exact-fp32 FFMA path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(M, K, N, seed):
    rng = np.random.RandomState(seed)
    x = rng.randn(M, K).astype(np.float32)
    w = (rng.randn(K, N) / np.sqrt(K)).astype(np.float32)
    sc, sh = (rng.rand(N) + 0.5).astype(np.float32), rng.randn(N).astype(np.float32)
    return x, w, sc, sh


@pytest.mark.timeout(120)
@pytest.mark.parametrize("M,K,N,act", [(128, 32, 128, 0), (256, 64, 128, 0), (1000, 256, 1024, 1), (4096, 192, 128, 1),
                                       (777, 128, 64, 1), (512, 64, 16, 1), (512, 16, 64, 2), (129, 1024, 256, 0),
                                       (8192, 512, 128, 1), (300, 36, 24, 0), (262144, 128, 64, 1)])
def test_linear_packed_vs_fp64(M, K, N, act):
    from dh3d_b200 import ops
    x, w, sc, sh = _case(M, K, N, M + K + N)
    dx, dw, dsc, dsh = (torch.from_numpy(a).cuda() for a in (x, w, sc, sh))
    packed = ops.linear_prepack(dw)
    y = ops.linear(dx, dw, scale=dsc, shift=dsh, act=act, packed=packed)
    torch.cuda.synchronize()
    if M * K * N <= 4e9:
        e = x.astype(np.float64) @ w.astype(np.float64) * sc + sh
    else:  # too slow in numpy fp64: compare against the exact-fp32 FFMA kernel instead
        e = ops.linear(dx, dw, scale=dsc, shift=dsh, act=0).cpu().numpy().astype(np.float64)
    e = np.maximum(e, 0) if act == 1 else (1 / (1 + np.exp(-e)) if act == 2 else e)
    err = np.abs(y.cpu().numpy() - e).max() / np.sqrt((e ** 2).mean())
    # measured ~3e-5 worst case (truncating fp32 accumulation inside the tensor core); bar 1e-4;
    # single-pass TF32 would give ~1e-3
    assert err < 6e-5, err


@pytest.mark.timeout(120)
def test_linear_packed_strided_output_and_exact_small_values():
    from dh3d_b200 import ops
    x, w, sc, sh = _case(640, 128, 128, 3)
    dx, dw = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    packed = ops.linear_prepack(dw)
    out = torch.full((640, 320), -7.0, device="cuda")
    ops.linear(dx, dw, out=out, out_col=64, packed=packed)
    ref = ops.linear(dx, dw)
    assert torch.all(out[:, :64] == -7) and torch.all(out[:, 192:] == -7)
    assert (out[:, 64:192] - ref).abs().max() <= 2e-5 * ref.pow(2).mean().sqrt()
    # integers up to 2^10 are tf32-exact: the product must be exact
    xi = torch.randint(-8, 9, (256, 64), device="cuda").float()
    wi = torch.randint(-8, 9, (64, 128), device="cuda").float()
    yi = ops.linear(xi, wi, packed=ops.linear_prepack(wi))
    assert torch.equal(yi, xi @ wi)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("M,K,N", [(1000, 256, 1024), (262144, 128, 64), (333, 64, 16), (4096, 512, 256)])
def test_linear_rowdot_fused_head(M, K, N):
    from dh3d_b200 import ops
    x, w, sc, sh = _case(M, K, N, 7 * M + N)
    rng = np.random.RandomState(N)
    w2 = (rng.randn(N) / np.sqrt(N)).astype(np.float32)
    dx, dw, dsc, dsh, dw2 = (torch.from_numpy(a).cuda() for a in (x, w, sc, sh, w2))
    y = ops.linear_rowdot(dx, ops.linear_prepack(dw), dsc, dsh, 1, dw2, 0.125, 2)
    if M * K * N <= 4e9:
        h = np.maximum(x.astype(np.float64) @ w.astype(np.float64) * sc + sh, 0)
        e = 1 / (1 + np.exp(-(h @ w2.astype(np.float64) + 0.125)))
    else:
        e = ops.rowdot(ops.linear(dx, dw, scale=dsc, shift=dsh, act=1), dw2, bias=0.125, act=2).cpu().numpy()
    assert np.abs(y.cpu().numpy() - e).max() < 5e-5
