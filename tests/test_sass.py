"""Static checks on the compiled kernels (cuobjdump -sass of the in-tree objects; no GPU needed).

Two things are pinned here that no numerical test on another shape would catch early:
  * rounding sequences that define bit-exact parity must survive the compiler: the 3-NN metric and the
    3-point interpolation are UN-fused in the reference (host code, tf_interpolate.cpp:60-127), so their
    kernels may not contain a single FFMA -- ptxas does contract `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2
    (seen while packing the k-NN scan), which is why the 3-NN metric is kept scalar;
  * the hot kernels really are sm_100a code: tcgen05 MMAs (UTCHMMA) with TMEM loads (LDTM) fed by TMA
    (UTMALDG / UTMASTG), cp.async staging (LDGSTS) and packed fp32x2 arithmetic (FFMA2) where DESIGN.md says so.
"""
import functools
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "dh3d_b200", "build")


def _cuobjdump():
    for c in (shutil.which("cuobjdump"), "/usr/local/cuda/bin/cuobjdump"):
        if c and os.path.exists(c):
            return c
    return None


@functools.lru_cache(maxsize=None)
def _functions(obj):
    """{mangled kernel name: [sass mnemonics]} of one object file."""
    exe = _cuobjdump()
    path = os.path.join(OBJ, obj)
    if exe is None or not os.path.exists(path):
        pytest.skip("cuobjdump or %s not available (run python -m dh3d_b200.build)" % obj)
    text = subprocess.run([exe, "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur.append(m.group(1))
    return out


def _count(ops, prefix):
    return sum(1 for o in ops if o == prefix or o.startswith(prefix + "."))


def _pick(funcs, *needles):
    hits = [k for k in funcs if all(n in k for n in needles)]
    assert hits, "no kernel matching %s in %s" % (needles, sorted(funcs)[:5])
    return hits


def test_unfused_metrics_contain_no_fma():
    knn = _functions("knn.o")
    for name in _pick(knn, "knn_query_kernel", "ThreeNnMetric"):
        assert _count(knn[name], "FFMA") == 0 and _count(knn[name], "FFMA2") == 0, name
        assert _count(knn[name], "FMUL") > 0 and _count(knn[name], "FADD") > 0, name
    gather = _functions("gather.o")
    for name in (_pick(gather, "three_interp_warp_kernelILb0E") + _pick(gather, "three_interp_kernelILi1ELb0E") +
                 _pick(gather, "three_nn_kernel")):
        assert _count(gather[name], "FFMA") == 0 and _count(gather[name], "FFMA2") == 0, name


def test_knn_and_fps_use_packed_fp32x2_with_explicit_fma_chains():
    knn = _functions("knn.o")
    for name in _pick(knn, "knn_query_kernelILi8ELb1ELb1E", "KnnMetric"):
        assert _count(knn[name], "FFMA2") >= 32 and _count(knn[name], "FMUL2") >= 16, name   # 32-point chunk, 16 pairs
    fps = _functions("fps.o")
    for name in _pick(fps, "fps_cluster_kernelILi8E"):
        assert _count(fps[name], "FFMA2") >= 8 and _count(fps[name], "FMUL2") >= 4, name     # 8 points = 4 pairs / round


def test_tensor_core_kernels_are_tcgen05_tma_code():
    gemm = _functions("gemm_tc16.o")
    for name in _pick(gemm, "gemm_tc16_kernelILi256ELb1ELi2E") + _pick(gemm, "gemm_join16_kernel"):
        ops = gemm[name]
        assert _count(ops, "UTCHMMA") >= 6, name          # tcgen05.mma, 3 split terms x 2 k-steps per slab
        assert _count(ops, "UTMALDG") >= 3, name          # TMA loads of the X / W tiles
        assert _count(ops, "LDTM") >= 1, name             # tcgen05.ld of the accumulator
    assert _count(gemm[_pick(gemm, "gemm_join16_kernel")[0]], "UTMASTG") >= 1     # TMA stores of y / its normalised copy
    head = _functions("gemm_head16.o")
    ops = head[_pick(head, "gemm_head16_kernel")[0]]      # resident-activation head kernel, CTA pairs
    assert _count(ops, "UTCHMMA.2CTA") >= 6 and _count(ops, "UTMALDG") >= 3 and _count(ops, "LDTM") >= 1
    assert _count(ops, "UTMASTG") == 0                    # the [M, 1024] hidden layer never leaves the SM
    pair = gemm[_pick(gemm, "gemm_tc16_kernelILi256ELb1ELi2E")[0]]
    assert _count(pair, "UTCHMMA.2CTA") >= 6              # tcgen05.mma.cta_group::2
    flex = _functions("flexconv_ca.o")
    for nb in (1, 2, 4):                                  # K = 8 (DH3D), 16, 32: the unrolled 8-slot schedule
        for name in (_pick(flex, "flexconv_ca_kernelILi64ELi%dE" % nb) + _pick(flex, "flexconv_ca_kernelILi128ELi%dE" % nb)):
            ops = flex[name]
            assert _count(ops, "UTCHMMA") >= 12 and _count(ops, "LDGSTS") >= 16 and _count(ops, "UTMASTG") >= 1, name
            assert _count(ops, "FFMA2") >= 8 * 12 and _count(ops, "FADD2") >= 8 * 4, name   # 8 unrolled neighbour slots
    nv = _functions("netvlad_tc.o")
    ops = nv[_pick(nv, "netvlad_tc2_kernel")[0]]
    assert _count(ops, "UTCHMMA") >= 72 and _count(ops, "UTMALDG") >= 8 and _count(ops, "LDTM") >= 2


def test_no_kernel_spills_to_local_memory_heavily():
    """Register budgets are part of the design (80 registers at 704 threads, 168 at 320): a change that pushes a
    hot loop into local memory shows up here as a jump in LDL/STL counts."""
    budgets = {("flexconv_ca.o", "flexconv_ca_kernelILi64ELi1E"): 40, ("gemm_head16.o", "gemm_head16_kernel"): 0, ("gemm_tc16.o", "gemm_join16_kernel"): 4,
               ("se_fused.o", "se_pool_excite_kernelILi64E"): 0, ("knn.o", "knn_query_kernelILi8ELb1ELb1E"): 8,
               ("netvlad_tc.o", "netvlad_tc2_kernel"): 24}
    for (obj, needle), limit in budgets.items():
        funcs = _functions(obj)
        for name in _pick(funcs, needle):
            n = _count(funcs[name], "LDL") + _count(funcs[name], "STL")
            assert n <= limit, "%s: %d local-memory instructions (budget %d)" % (name, n, limit)
