"""The C-ABI library loads on a CPU-only box, exports every symbol include/orbfe.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from monoorbslam3_b200 import _capi, build
    if not os.path.exists(_capi.LIB_PATH):
        build.build()
    return _capi


def header_symbols(name="orbfe.h"):
    txt = open(os.path.join(ROOT, "include", name)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orbfe_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_the_reference_facing_entry_points():
    syms = header_symbols()
    for s in ("orbfe_create", "orbfe_extract", "orbfe_extract_batch", "orbfe_extract_batch_device", "orbfe_hamming_allpairs",
              "orbfe_descriptor_distance", "orbfe_search_for_initialization", "orbfe_search_by_projection", "orbfe_search_local_points",
              "orbfe_search_for_triangulation"):
        assert s in syms


def test_library_exports_every_declared_symbol(capi):
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert set(header_symbols()) == set(capi.SIGNATURES)          # the ctypes table covers the header, nothing more


def test_multi_gpu_library_exports_every_declared_symbol(capi):
    """include/orbfe_dist.h / liborbfe_dist.so: the multi-GPU entry points (NCCL from C++)."""
    from monoorbslam3_b200 import build
    out = subprocess.run(["nm", "-D", "--defined-only", build.build_dist()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (orbfe_[a-z_0-9]+)", out))
    declared = set(header_symbols("orbfe_dist.h"))
    assert declared and declared == exported, (declared ^ exported)
    assert {"orbfe_dist_init", "orbfe_extract_batch_sharded", "orbfe_extract_batch_sharded_device", "orbfe_allpairs_sharded"} <= declared


def test_library_contains_sm100a_code_only(capi):
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(5|6|7|8|9)\d", out)


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.OrbfeError) as e:
        capi.create(1000, 1.2, 8, 20, 7)
    assert e.value.code == capi.ORBFE_E_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "monoorbslam3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert "oracle" not in open(os.path.join(dirpath, f), errors="ignore").read().replace("ORBFE", ""), os.path.join(dirpath, f)


def test_built_kernels_are_blackwell_native():
    """The shipped liborbfe.so holds sm_100a code whose all-pairs kernel issues tcgen05 MMAs with the accumulators in tensor memory
    (UTCIMMA, LDTM) on TMA-staged operands (UTMALDG), and the extractor's tiles come in by TMA as well (cuobjdump -sass, no GPU needed)."""
    import shutil
    import subprocess
    from monoorbslam3_b200 import build as _b
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    lib = _b.build()

    def sass(fn):
        return subprocess.run([exe, "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout

    tc = sass("_ZN5orbfe13k_allpairs_tcE14CUtensorMap_stS0_NS_6TcArgsE")
    assert "sm_100a" in tc
    for mnemonic, at_least in (("UTCIMMA", 16), ("LDTM", 4), ("UTMALDG", 2), ("UTCBAR", 1)):
        assert tc.count(mnemonic) >= at_least, (mnemonic, tc.count(mnemonic))
    fast = sass("_ZN5orbfe13k_fast_planesILb1EEEvNS_8LevelSetENS_7TmapSetENS_9Fast2ArgsE")
    assert fast.count("UTMALDG") >= 1 and fast.count("IDP.4A") >= 32 and fast.count("VABSDIFF4") >= 32
