"""Compile the REFERENCE's own sources, unmodified and where they lie under /root/reference,
into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).  Test/bench
infrastructure only -- see oracle/__init__.py.

  libdh3d_ref_cuda.so  user_ops/kernels/{knn_bruteforce,flex_conv,flex_pool,conv_pointset,flex_deconv}_kernel_gpu.cu.cc
                       + tf_ops/sampling/tf_sampling_g.cu + tf_ops/grouping/tf_grouping_g.cu, built
                       for sm_100a with the reference's own flags (-O3 / -O2, no fast-math;
                       user_ops/CMakeLists.txt:32, tf_ops/*/tf_*_compile.sh) against a stub of the two
                       TensorFlow headers they include (oracle/ref_shim/stub), CUDA 12.9's bundled CUB.
  libdh3d_ref_cpu.so   user_ops/kernels/{flex_conv,flex_pool,conv_pointset,flex_deconv}_kernel.cc (CPU functors,
                       forward and Grad)
                       + the plain-C functions of tf_ops/interpolation/tf_interpolate.cpp (lines
                       55-153 are extracted at BUILD time into oracle/_ref/; that file also holds
                       TF op registrations that cannot compile without TensorFlow), g++ -O2, no -mfma
                       like tf_interpolate_compile.sh:11-15.
  dh3d_weights.npz     the MODEL VARIABLES of the two shipped checkpoints (models/{local,global}; optimizer
                       slots dropped), keys 'local:<tf name>' / 'global:<tf name>': data, not source -- lets the
                       real-weight parity tests run on the GPU box, where /root/reference does not exist.
The reference's own build system (cmake + FindTensorFlow) is not run: TensorFlow is absent.
The k-NN CPU functor needs Eigen (absent) and is not built.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DH3D_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim")
STUB = os.path.join(SHIM, "stub")
CUDA_SO = os.path.join(OUT, "libdh3d_ref_cuda.so")
CPU_SO = os.path.join(OUT, "libdh3d_ref_cpu.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), res.stdout[-3000:], res.stderr[-3000:]))


def _gxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _fresh(so, shim_src):
    """The prebuilt library is kept unless the shim that defines its entry points is newer."""
    return os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(os.path.join(SHIM, shim_src))


def build_cuda(force=False):
    if _fresh(CUDA_SO, "ref_cuda_shim.cu") and not force:
        return CUDA_SO
    kern = os.path.join(REF, "user_ops", "kernels")
    objs = []
    for name in ("knn_bruteforce", "flex_conv", "flex_pool", "conv_pointset", "flex_deconv"):
        obj = os.path.join(OUT, name + "_gpu.o")
        _run(["nvcc"] + ARCH + ["-O3", "-std=c++17", "--expt-relaxed-constexpr", "-DGOOGLE_CUDA=1", "-w",
                                "-I", STUB, "-I", kern, "-Xcompiler", "-fPIC", "-x", "cu", "-c",
                                os.path.join(kern, name + "_kernel_gpu.cu.cc"), "-o", obj])
        objs.append(obj)
    for rel in ("tf_ops/sampling/tf_sampling_g.cu", "tf_ops/grouping/tf_grouping_g.cu"):
        obj = os.path.join(OUT, os.path.basename(rel) + ".o")
        _run(["nvcc"] + ARCH + ["-O2", "-DGOOGLE_CUDA=1", "-w", "-Xcompiler", "-fPIC", "-x", "cu", "-c",
                                os.path.join(REF, rel), "-o", obj])
        objs.append(obj)
    shim = os.path.join(OUT, "ref_cuda_shim.o")
    _run(["nvcc"] + ARCH + ["-O2", "-std=c++17", "-DGOOGLE_CUDA=1", "-w", "-I", STUB, "-I", kern,
                            "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-c",
                            os.path.join(SHIM, "ref_cuda_shim.cu"), "-o", shim])
    _run(["nvcc", "-shared"] + ARCH + ["-o", CUDA_SO, shim] + objs)
    return CUDA_SO


def build_cpu(force=False):
    if _fresh(CPU_SO, "ref_cpu_shim.cpp") and not force:
        return CPU_SO
    kern = os.path.join(REF, "user_ops", "kernels")
    # the plain-C functions of tf_interpolate.cpp (threenn_cpu .. threeinterpolate_grad_cpu)
    src = open(os.path.join(REF, "tf_ops", "interpolation", "tf_interpolate.cpp")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("void threenn_cpu"))
    end = next(i for i, l in enumerate(src) if l.startswith("class ThreeNNOp"))
    extracted = os.path.join(OUT, "tf_interpolate_fns.cpp")
    with open(extracted, "w") as f:
        f.write("// extracted at build time from /root/reference/tf_ops/interpolation/tf_interpolate.cpp"
                " lines %d-%d; not committed\n#include <cmath>\n#include <cstring>\n" % (start + 1, end))
        f.write("\n".join(src[start:end]))
    objs = []
    o = os.path.join(OUT, "tf_interpolate_fns.o")
    _run([_gxx(), "-std=c++11", "-O2", "-fPIC", "-c", extracted, "-o", o])
    objs.append(o)
    for name in ("flex_conv", "flex_pool", "conv_pointset", "flex_deconv"):
        o = os.path.join(OUT, name + "_cpu.o")
        _run([_gxx(), "-std=c++14", "-O3", "-fPIC", "-w", "-I", STUB, "-I", kern, "-c",
              os.path.join(kern, name + "_kernel.cc"), "-o", o])
        objs.append(o)
    shim = os.path.join(OUT, "ref_cpu_shim.o")
    _run([_gxx(), "-std=c++14", "-O2", "-fPIC", "-fvisibility=hidden", "-w", "-I", STUB, "-I", kern, "-c",
          os.path.join(SHIM, "ref_cpu_shim.cpp"), "-o", shim])
    _run([_gxx(), "-shared", "-o", CPU_SO, shim] + objs)
    return CPU_SO


WEIGHTS_NPZ = os.path.join(OUT, "dh3d_weights.npz")


def stage_weights(force=False):
    if os.path.exists(WEIGHTS_NPZ) and not force:
        return WEIGHTS_NPZ
    import numpy as np
    sys.path.insert(0, os.path.dirname(HERE))
    from dh3d_b200.checkpoint import read_tensor_bundle
    out = {}
    for tag in ("local", "global"):
        for k, v in read_tensor_bundle(os.path.join(REF, "models", tag, tag + "model")).items():
            if "Adam" in k or k.startswith(("EMA/", "beta")) or k in ("global_step", "learning_rate"):
                continue
            out["%s:%s" % (tag, k)] = np.asarray(v)
    np.savez(WEIGHTS_NPZ, **out)
    return WEIGHTS_NPZ


def load_staged_weights():
    """-> ({tf_name: ndarray} local, {tf_name: ndarray} global) or None when not staged."""
    if not os.path.exists(WEIGHTS_NPZ):
        return None
    import numpy as np
    z = np.load(WEIGHTS_NPZ)
    loc = {k[6:]: z[k] for k in z.files if k.startswith("local:")}
    glo = {k[7:]: z[k] for k in z.files if k.startswith("global:")}
    return loc, glo


def build(force=False, verbose=False):
    """No-op (returns None) when /root/reference is absent, e.g. on the GPU box."""
    if not os.path.isdir(os.path.join(REF, "user_ops", "kernels")):
        if verbose:
            print("reference tree not present: keeping prebuilt oracle/_ref (if any)")
        return None
    os.makedirs(OUT, exist_ok=True)
    a, b = build_cuda(force), build_cpu(force)
    if os.path.isdir(os.path.join(REF, "models", "global")):
        stage_weights(force)
    if verbose:
        print("built", a, "and", b)
    return a, b


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
