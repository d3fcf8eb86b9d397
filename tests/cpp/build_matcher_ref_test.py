"""g++ build of tests/cpp/matcher_ref_test.cpp (test infrastructure): the product's matcher adapter (monoorbslam3_b200/host, compiled with
ORBFE_REFERENCE_TYPES) against the stand-in Frame / KeyFrame / MapPoint headers of oracle/matchshim and the reference's vendored DBoW2
FeatureVector, which is compiled from where it lies under /root/reference — so the binary is built in the build container
(__graft_entry__.build()) and travels to the GPU box prebuilt (tests/cpp/_build is git-ignored, not gpurun-ignored)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build(out=None):
    out = out or os.path.join(HERE, "_build", "matcher_ref_test")
    pkg = os.path.join(ROOT, "monoorbslam3_b200")
    lib = os.path.join(pkg, "lib", "liborbfe.so")
    ref = os.environ.get("ORBFE_REFERENCE", "/root/reference")
    dbow = os.path.join(ref, "thirdParty", "DBoW2")
    if not os.path.exists(os.path.join(dbow, "DBoW2", "FeatureVector.cpp")):
        if os.path.exists(out):
            return out
        raise FileNotFoundError("matcher_ref_test is not built and %s is not present" % dbow)
    host, orc = os.path.join(pkg, "host"), os.path.join(ROOT, "oracle")
    srcs = [os.path.join(HERE, "matcher_ref_test.cpp"), os.path.join(host, "ORBExtractor.cpp"), os.path.join(host, "ORBMatcher.cpp"),
            os.path.join(dbow, "DBoW2", "FeatureVector.cpp"), os.path.join(dbow, "DBoW2", "BowVector.cpp")]
    deps = srcs + [lib, os.path.join(host, "ORBMatcher.h"), os.path.join(host, "ORBExtractor.h"), os.path.join(orc, "matchshim", "BasicObject", "Frame.h")]
    if os.path.exists(out) and all(os.path.getmtime(s) < os.path.getmtime(out) for s in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["make", "-s", "-C", orc, os.path.join(orc, "liborb_oracle.so")])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-DORBFE_REFERENCE_TYPES", "-I" + host, "-I" + os.path.join(orc, "matchshim"),
                           "-I" + os.path.join(orc, "eigenshim"), "-I" + os.path.join(orc, "cvshim"), "-I" + os.path.join(orc, "boostshim"), "-I" + orc,
                           "-I" + dbow, "-o", out] + srcs +
                          ["-L" + os.path.dirname(lib), "-lorbfe", "-L" + orc, "-lorb_oracle", "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + orc])
    return out


if __name__ == "__main__":
    print(build())
