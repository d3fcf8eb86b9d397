"""Build liborbfe.so (CUDA kernels + C-ABI) in-tree for sm_100a with nvcc.  Called by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "liborbfe.so")
SOURCES = ["orbfe_extract.cu", "orbfe_match.cu", "orbfe_allpairs_tc.cu", "orbfe_frame.cu", "orbfe_bow.cu", "orbfe_geom.cu"]
HEADERS = ["orbfe_internal.cuh", "orbfe_kernels.cuh", "brief_pattern.inc", os.path.join("..", "..", "include", "orbfe.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-shared", "-cudart", "static"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd)
    return LIB


DIST_LIB = os.path.join(HERE, "lib", "liborbfe_dist.so")


def build_dist(force=False):
    """liborbfe_dist.so: the multi-GPU entry points (include/orbfe_dist.h) — host C++ over liborbfe.so and NCCL."""
    src = os.path.join(CSRC, "orbfe_dist.cpp")
    deps = [src, LIB, os.path.join(HERE, "..", "include", "orbfe_dist.h"), os.path.join(HERE, "..", "include", "orbfe.h")]
    if not force and os.path.exists(DIST_LIB) and all(os.path.getmtime(d) < os.path.getmtime(DIST_LIB) for d in deps):
        return DIST_LIB
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I" + os.path.join(cuda, "include"), "-o", DIST_LIB, src,
                           "-L" + os.path.dirname(LIB), "-lorbfe", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lnccl", "-lpthread",
                           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    return DIST_LIB


def build_cpp_dist_test(out=None):
    """g++ build of tests/cpp/dist_test.cpp against liborbfe_dist.so / liborbfe.so (used by the multi-GPU test)."""
    root = os.path.dirname(HERE)
    out = out or os.path.join(HERE, "lib", "dist_test")
    src = os.path.join(root, "tests", "cpp", "dist_test.cpp")
    build_dist()
    if os.path.exists(out) and all(os.path.getmtime(s) < os.path.getmtime(out) for s in (src, LIB, DIST_LIB)):
        return out
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(cuda, "include"), "-o", out, src, "-L" + os.path.dirname(LIB), "-lorbfe_dist", "-lorbfe",
                           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.dirname(LIB), "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_dist(force="--force" in sys.argv))


def build_cpp_adapter_test(out=None):
    """g++ build of the C++ host adapters + tests/cpp/adapter_test.cpp against liborbfe.so (used by the GPU test-suite)."""
    root = os.path.dirname(HERE)
    out = out or os.path.join(HERE, "lib", "adapter_test")
    srcs = [os.path.join(HERE, "host", "ORBExtractor.cpp"), os.path.join(HERE, "host", "ORBMatcher.cpp"), os.path.join(root, "tests", "cpp", "adapter_test.cpp")]
    deps = srcs + [LIB] + [os.path.join(HERE, "host", f) for f in ("ORBExtractor.h", "ORBMatcher.h", "FramePost.h", "cv_compat.h")]
    if os.path.exists(out) and all(os.path.getmtime(s) < os.path.getmtime(out) for s in deps):
        return out
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", out] + srcs + ["-L" + os.path.dirname(LIB), "-lorbfe", "-Wl,-rpath," + os.path.dirname(LIB)])
    return out

